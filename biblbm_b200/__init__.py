"""biblbm_b200 -- B200-native (sm_100a CUDA + NCCL) hot path of niclasberg/biblbm behind a C-ABI.

Only what the path needs lives here: csrc/ (CUDA kernels + the C-ABI of include/biblbm_b200.h), build.py (nvcc build,
in-tree) and api.py (ctypes host mirror of the reference's lattice / immersed-boundary interface).
"""
from .api import (BibError, Lattice, comm_unique_id, device_count, load_library, set_default_inplace, inplace_schedule, SYMBOLS,  # noqa: F401
                  OPT_ENVELOPE_SWAP, OPT_DETERMINISTIC_SPREAD, OPT_REGULATOR, OPT_SPREAD_MODE, OPT_SPARSE_FORCE, OPT_STEP_GRAPH, OPT_FUSED_FACE_SWAP, OPT_OVERLAP_FORCES, OPT_PREFETCH_FORCES, OPT_FUSED_INTERPOLATION, OPT_FUSE_CALLS, OPT_LBM_PREFETCH, OPT_INTERP_COMPACT, CELL_BULK, CELL_NODYNAMICS, CELL_BOUNCEBACK, PARTICLE_RBC, PARTICLE_CAPSULE, PARTICLE_SEMIRIGID,
                  POTENTIAL_NONE, POTENTIAL_MORSE, POTENTIAL_SPRING)
