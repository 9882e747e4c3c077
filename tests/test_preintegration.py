"""Leg-contact preintegration: oracle self-consistency and the device kernel (through the CPU kernel simulator)."""
import ctypes as C
import numpy as np
import pytest
from cerberus_b200 import abi, synth
from oracle_lib import OracleBackend
from helpers import sim_backend, small_cfg

ob = OracleBackend()


def make_jobs(n=3, seed=0, contact_type=0):
    batch, truth = synth.generate_batch(1, 4, ob, with_prior=False, return_truth=True, window0=seed)
    jobs = (abi.PreintJob * n)()
    for k in range(n):
        C.memmove(C.byref(jobs[k]), C.byref(truth.raw_jobs[k]), C.sizeof(abi.PreintJob))
    if contact_type == 2:   # foot force readings instead of 0/1 flags
        for k in range(n):
            s = np.ctypeslib.as_array(C.cast(jobs[k].samples, C.POINTER(C.c_double)), shape=(jobs[k].n_samples, 35))
            s[:, 31:35] = 80.0 * s[:, 31:35] + 5.0
    jobs._keepalive = (truth, batch)      # the jobs point into truth.raw_samples
    return jobs, truth, batch


def test_bias_jacobian_predicts_repropagation():
    """jacobian (d delta / d bias) is a first-order model of re-integrating with a different bias, the property
    IntegrationBase::checkJacobian of the reference prints (integration_base.h:292-458)."""
    jobs, truth, _ = make_jobs(1)
    pcfg = abi.default_preint_config()
    base = ob.preintegrate(pcfg, jobs, 1)[0]
    J = np.array(base["jacobian"]).reshape(31, 31).T          # column-major -> J[r, c]
    dbg = np.array([2e-4, -1e-4, 3e-4]); dba = np.array([1e-3, 2e-3, -1e-3]); drho = np.array([1e-4, -2e-4, 1e-4, 2e-4])
    j2 = (abi.PreintJob * 1)(); C.memmove(C.byref(j2[0]), C.byref(jobs[0]), C.sizeof(abi.PreintJob))
    for k in range(3):
        j2[0].linearized_ba[k] += dba[k]; j2[0].linearized_bg[k] += dbg[k]
    for k in range(4):
        j2[0].linearized_rho[k] += drho[k]
    pert = ob.preintegrate(pcfg, j2, 1)[0]
    dp = np.array(pert["delta_p"]) - np.array(base["delta_p"])
    assert np.abs(dp - (J[0:3, 21:24] @ dba + J[0:3, 24:27] @ dbg)).max() < 2e-3 * np.abs(dp).max() + 1e-9
    dv = np.array(pert["delta_v"]) - np.array(base["delta_v"])
    assert np.abs(dv - (J[6:9, 21:24] @ dba + J[6:9, 24:27] @ dbg)).max() < 2e-3 * np.abs(dv).max() + 1e-9
    de = (np.array(pert["delta_epsilon"]) - np.array(base["delta_epsilon"])).reshape(4, 3)
    for leg in range(4):
        pred = J[9 + 3 * leg:12 + 3 * leg, 24:27] @ dbg + J[9 + 3 * leg:12 + 3 * leg, 27 + leg] * drho[leg]
        assert np.abs(de[leg] - pred).max() < 5e-3 * np.abs(de[leg]).max() + 1e-9


def test_covariance_is_symmetric_psd_and_contact_dependent():
    jobs, _, _ = make_jobs(2)
    out = ob.preintegrate(abi.default_preint_config(), jobs, 2)
    for k in range(2):
        cov = np.array(out[k]["covariance"]).reshape(31, 31)
        assert np.abs(cov - cov.T).max() <= 1e-12 * np.abs(cov).max()
        assert np.linalg.eigvalsh(0.5 * (cov + cov.T)).min() > 0
        # epsilon variance is dominated by V_N_MAX * dt^2 per swing sample (imu_leg_integration_base.cpp:290-299,460):
        # legs of the two trot pairs spend a different number of samples in the air inside one interval
        d = np.diag(cov)[9:21].reshape(4, 3)
        dt = 1.0 / 15 / 33
        nswing = d[:, 0] / (900.0 * dt * dt)
        assert np.abs(nswing - np.round(nswing)).max() < 0.05 and (nswing > 0).all() and nswing.max() <= 33.01


@pytest.mark.parametrize("contact_type", [0, 2])
def test_device_preintegration_matches_oracle(contact_type):
    jobs, _, _ = make_jobs(3, seed=5, contact_type=contact_type)
    pcfg = abi.default_preint_config(); pcfg.contact_sensor_type = contact_type
    ref = ob.preintegrate(pcfg, jobs, 3)
    got = sim_backend(small_cfg()).preintegrate(pcfg, jobs, 3)
    for name in ref.dtype.names:
        scale = max(1e-30, np.abs(ref[name]).max())
        assert np.abs(got[name] - ref[name]).max() / scale < 1e-12, name
