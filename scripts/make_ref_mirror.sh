#!/bin/bash
# Writable mirror of the reference's UNCHANGED FEMShell scripts, Python drivers and input meshes under baseline/_ref/
# (git-ignored, not gpurun-ignored: it travels to the GPU box, the reference tree itself does not exist there).
# Nothing of it is committed. Usage: scripts/make_ref_mirror.sh [/root/reference]
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
DST=$HERE/baseline/_ref/IDP_mirror
[ -d "$REF/Projects/FEMShell" ] || { echo "no reference tree at $REF"; exit 0; }
mkdir -p "$DST/Projects/FEMShell/input" "$DST/Python"
cp "$REF"/Projects/FEMShell/*.py "$DST/Projects/FEMShell/"
cp -r "$REF/Python/Drivers" "$DST/Python/"
for m in bunny3K hand cat feline font_Tao wm2_15k; do
  [ -f "$REF/Projects/FEMShell/input/$m.obj" ] && cp "$REF/Projects/FEMShell/input/$m.obj" "$DST/Projects/FEMShell/input/"
done
# config 2 (16_fix_char_seq.py): rest mannequin + the first target frames of the Rumba sequence
for seq in Rumba_Dancing_unfixed; do
  mkdir -p "$DST/Projects/FEMShell/input/$seq"
  for f in $(seq 0 ${MIRROR_FRAMES:-6}); do
    [ -f "$REF/Projects/FEMShell/input/$seq/$f.obj" ] && cp "$REF/Projects/FEMShell/input/$seq/$f.obj" "$DST/Projects/FEMShell/input/$seq/"
  done
done
# batch.py's second sequence: three target frames
mkdir -p "$DST/Projects/FEMShell/input/Kick_unfixed"
for f in 0 1 2 3; do
  [ -f "$REF/Projects/FEMShell/input/Kick_unfixed/$f.obj" ] && cp "$REF/Projects/FEMShell/input/Kick_unfixed/$f.obj" "$DST/Projects/FEMShell/input/Kick_unfixed/"
done
echo "mirror at $DST"
