"""Marginalization half of Estimator::optimization() (reference src/estimator/estimator.cpp:1247-1456 +
src/factor/marginalization_factor.cpp:12-333): a thin host wrapper over cerb_batch_marginalize.

Everything numerical runs on the device (csrc/solve_kernel.cuh marg_assemble_kernel + csrc/marg_kernels.cuh): the factors that touch the
dropped blocks are linearised by the solve kernel's own visual / inertial passes (same Huber corrector), A = sum J^T J and b = sum J^T r are
assembled in the reference's [dropped | kept] order and reduced by the eps = 1e-8 clamped eigen Schur complement.  What is left here is
moving the result into the next window's descriptor.

Block order inside the new prior is fixed (the reference's order is that of an unordered_map keyed by pointer value, i.e. arbitrary):
dropped = [pose0, speedbias0, legbias0, lambdas...], kept = [pose k ..., speedbias1, legbias1, ex0, ex1, td] restricted to the blocks
that occur.
"""
import ctypes as C
import numpy as np
from . import abi


def marginalize_batch(backend, cfg, src, dst, margin_old=True, upload=True):
    """Marginalize every window of `src` at its current host states and write the resulting prior, already address-shifted for the next
    window, into the descriptors of `dst` (dst may be src).  margin_old: one flag for the batch or one per window (True: MARGIN_OLD).
    upload=False: `src`'s descriptors are already resident on the device (e.g. right after solve_batch on the same batch)."""
    B = src.n
    flags = np.ascontiguousarray(np.where(np.broadcast_to(np.asarray(margin_old, dtype=bool), (B,)), 0, 1), dtype=np.int32)
    if upload:
        backend.upload(src)
    J = np.zeros((B, abi.MAX_PRIOR_DIM * abi.MAX_PRIOR_DIM)); r = np.zeros((B, abi.MAX_PRIOR_DIM))
    priors = (abi.Prior * B)()
    for w in range(B):
        priors[w].linearized_jacobians = J[w].ctypes.data_as(abi.c_dp); priors[w].linearized_residuals = r[w].ctypes.data_as(abi.c_dp)
    sweeps = backend.batch_marginalize(flags, src.states, priors)
    for w in range(B):
        pd = dst.descs[w].prior
        keepJ, keepr = dst.prior_J[w].ctypes.data_as(abi.c_dp), dst.prior_r[w].ctypes.data_as(abi.c_dp)
        C.memmove(C.byref(pd), C.byref(priors[w]), C.sizeof(abi.Prior))
        dst.prior_J[w] = J[w]; dst.prior_r[w] = r[w]
        pd.linearized_jacobians, pd.linearized_residuals = keepJ, keepr
    return sweeps
