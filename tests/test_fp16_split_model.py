"""CPU model of the split-FP16 arithmetic of the tensor-core power flow (csrc/powerflow_tc2.cu):
operands scaled by powers of two into the FP16 range, x = x_hi + x_lo, B = B_hi + B_lo, three
products accumulated in FP32.  Checks the accuracy claim of DESIGN.md 4.3b on the feeders the
repository ships, and the feeder property the kernel's node expansion relies on (a wye load's
node voltage is its branch voltage times the ratio of the voltage bases)."""
import numpy as np
import pytest

from powergridworld_b200.distribution_system.feeder import compile_feeder

FEEDERS = ["synthetic123.dss", "ieee_13_dss/IEEE13Nodeckt.dss"]


def _realify(z):
    return np.block([[-z.real, z.imag], [-z.imag, -z.real]])


def _split16(a, scale):
    a = a * scale
    hi = a.astype(np.float16)
    lo = (a - hi.astype(np.float64)).astype(np.float32).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def _scale_for(m):
    return 2.0 ** np.floor(np.log2(16384.0 / np.abs(m).max()))


def _converged_currents(f, load_scale=1.0):
    s = (f.load_kw + 1j * f.load_kvar)[f.branch_load] * f.branch_share * 1e-3 * load_scale
    u = f.u0.copy()
    for _ in range(60):
        un = f.u0 - f.zbb @ np.conj(s / u)
        if np.abs(un - u).max() < 1e-12:
            break
        u = un
    return np.conj(s / u)


@pytest.mark.parametrize("name", FEEDERS)
@pytest.mark.parametrize("load_scale", [0.3, 1.0])
def test_three_product_split_is_accurate_to_1e7(name, load_scale):
    f = compile_feeder(name)
    i = _converged_currents(f, load_scale)
    x = np.concatenate([i.real, i.imag])
    xh, xl = _split16(x, 2048.0)
    assert np.abs(xh).max() < 65504, "current scale leaves the FP16 range"
    for z in (f.zbb, f.znb):
        b = _realify(z)
        sb = _scale_for(b)
        bh, bl = _split16(b, sb)
        assert np.abs(bh).max() < 65504
        acc = (bh @ xl).astype(np.float32) + (bl @ xh).astype(np.float32)
        acc = (acc + (bh @ xh).astype(np.float32)).astype(np.float64) / (sb * 2048.0)
        exact = b @ x
        assert np.abs(acc - exact).max() < 1e-7, "three-product chain"
        single = (bh.astype(np.float64) @ xh.astype(np.float64)) / (sb * 2048.0)
        assert np.abs(single - exact).max() > 10 * np.abs(acc - exact).max(), \
            "the correction products are what buys the accuracy"


@pytest.mark.parametrize("name,expected", [(FEEDERS[0], 85), (FEEDERS[1], 11)])
def test_wye_load_nodes_derive_from_branch_voltages(name, expected):
    """Row n of Znb = c x row k of Zbb and w[n] = c u0[k] for every wye load branch k at node n;
    pgw_create detects exactly this to skip those nodes in the expansion."""
    f = compile_feeder(name)
    hits = 0
    for n in range(f.nn):
        for k in range(f.nb):
            c = (f.w[n] / f.u0[k]).real
            if c > 0 and abs(f.w[n] - c * f.u0[k]) <= 1e-10 * abs(f.w[n]) and \
                    np.abs(f.znb[n] - c * f.zbb[k]).max() <= 1e-10 * np.abs(f.znb[n]).max():
                hits += 1
                break
    assert hits == expected
