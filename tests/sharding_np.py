"""Host-side partition logic of the sharded path, mirrored from the CUDA sources so that it can be tested on CPU
(tests/test_sharding_gloo.py):

* query primitives: rank r of P owns the contiguous slice [n*r//P, n*(r+1)//P) (shard_range() in csrc/exact_kernels.cu);
* constraint rows built by idp_constraint_set stay on the rank that produced them (LOCAL-ROWS mode): its own direct PT /
  EE rows plus the merged PP / PE rows whose leading vertex lies in its vertex slab; the global list is only gathered on
  request (comm_gather_groups() in csrc/comm.cu: direct groups in rank order, merged group in descending rank order);
* rows handed in by the caller (idp_set_constraints) are replicated, and each is evaluated by the rank that owns the
  vertex chunk of its smallest vertex, chunks of 2^shift consecutive vertices dealt round-robin (row_owner() in
  csrc/barrier_kernels.cu).
Either way rows touching the same vertices land on the same rank, so the per-rank partial CSRs are nearly disjoint; their
sum is the global Hessian.
"""
import numpy as np


def shard_range(n, rank, nranks):
    return (n * rank) // nranks, (n * (rank + 1)) // nranks


def owner_shift(n_vertices, nranks):
    """chunk size exponent: about 16 chunks per rank, between 2^8 and 2^14 vertices"""
    s = 8
    while s < 14 and (n_vertices >> s) > 16 * nranks:
        s += 1
    return s


def row_vertices(rows):
    """(n,4) constraint rows -> (n,4) stencil vertices, -1 where unused (decode_row, csrc/pair_exact.cuh)"""
    r = np.asarray(rows, np.int64)
    v = np.where(r < 0, -r - 1, r)
    unused = np.zeros(r.shape, bool)
    dup = (r[:, 0] < 0) & (r[:, 3] < 0)           # merged PP / PE rows: slot 3 is -multiplicity
    unused[dup, 3] = True
    unused[dup & (r[:, 2] < 0), 2] = True        # PP: slot 2 is -1
    return np.where(unused, -1, v)


def row_owner(rows, n_vertices, nranks):
    v = row_vertices(rows)
    mv = np.where(v < 0, np.iinfo(np.int64).max, v).min(axis=1)
    return (mv >> owner_shift(n_vertices, nranks)) % nranks


def combine_energy(parts):
    """ncclAllReduce(sum) of the per-rank partial energies"""
    return float(sum(parts))


def combine_step(parts):
    """ncclAllReduce(min) of the per-rank CCD steps"""
    return float(min(parts))
