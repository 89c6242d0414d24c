"""idp_b200 — B200-native (sm_100a) IPC contact hot path of ipc-sim/IDP behind a C ABI (include/idp_contact.h).

Package contents: csrc/ (CUDA kernels + extern "C" launchers -> libidp_contact.so), host/ (C++ mirror of the
reference's six IPC.h operators), contact.py (ctypes harness for tests / bench), meshgen.py (synthetic inputs).
"""
from .contact import ContactContext, IdpError, load_library, LIB_PATH, EXPORTS  # noqa: F401
