"""GPU: the reference's six-function plugin ABI served by the shim libraries — the same call script
is run against libvdec_{awgn,bsc}.so and (when built) against the compiled reference objects."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT, make_hard_inputs, make_soft_inputs, ocode

pytestmark = pytest.mark.gpu

REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _targets(kind):
    t = [None]
    p = os.path.join(REFDIR, f"libref_{kind}.so")
    if os.path.exists(p):
        t.append(p)
    return t


@pytest.mark.parametrize("idx", [0, 3, 5])
def test_awgn_per_symbol_calls(cc, orc, idx):
    """AWGN driver pattern: one trellis step (M floats) per decoder_input call (main.c:108-116)"""
    code = cc.get_code(idx)
    _, met = make_soft_inputs(orc, ocode(code), 6, 0.5, seed=idx)
    exp = orc.decode_soft(ocode(code), met)
    for lib in _targets("awgn"):
        d = cc.LegacyDecoder("awgn", lib)
        assert d.init(code) == 0
        d.register_callback()
        for f in range(6):
            for t in range(code.T):
                before = len(d.frames)
                assert d.input(met[f, t]) == 0
                # the callback fires synchronously inside the call that delivers step T
                assert len(d.frames) == before + (1 if t == code.T - 1 else 0)
        got = np.frombuffer(b"".join(fr[0] for fr in d.frames), np.uint8).reshape(6, code.out_bytes)
        assert np.array_equal(got, exp) and all(fr[1] == code.L for fr in d.frames)
        d.destroy()
        assert not d.obj                                              # destroy NULLs the caller's pointer


def test_awgn_multi_symbol_calls_bad_len_reset_reinit(cc, orc):
    code = cc.get_code(0)
    _, met = make_soft_inputs(orc, ocode(code), 3, 0.5, seed=11)
    exp = orc.decode_soft(ocode(code), met)
    for lib in _targets("awgn"):
        d = cc.LegacyDecoder("awgn", lib)
        assert d.init(code) == 0
        d.register_callback()
        assert d.input(np.zeros(5, np.float32)) == -1                 # len % 2^n != 0 (viterbi-decoder.c:225-228)
        flat = met[0].reshape(-1)
        for off in range(0, flat.size, 28):                           # 7 steps per call
            assert d.input(flat[off:off + 28]) == 0
        assert len(d.frames) == 1
        # mid-frame reset aborts the frame (:212-222)
        assert d.input(met[1, :10].reshape(-1)) == 0
        d.reset()
        for t in range(code.T):
            d.input(met[1, t])
        assert len(d.frames) == 2
        # re-init with another code, callback kept
        code5 = cc.get_code(5)
        _, met5 = make_soft_inputs(orc, ocode(code5), 1, 0.5, seed=12)
        assert d.init(code5) == 0
        for t in range(code5.T):
            d.input(met5[0, t])
        assert len(d.frames) == 3
        assert np.array_equal(np.frombuffer(d.frames[0][0], np.uint8), exp[0])
        assert np.array_equal(np.frombuffer(d.frames[1][0], np.uint8), exp[1])
        assert np.array_equal(np.frombuffer(d.frames[2][0], np.uint8), orc.decode_soft(ocode(code5), met5)[0])
        d.destroy()


def test_awgn_null_callback_and_ownership_transfer(cc, orc):
    code = cc.get_code(0)
    _, met = make_soft_inputs(orc, ocode(code), 2, 0.5, seed=21)
    exp = orc.decode_soft(ocode(code), met)
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    for lib in _targets("awgn"):
        d = cc.LegacyDecoder("awgn", lib)
        assert d.init(code) == 0
        for t in range(code.T):                                       # no callback registered: treated as -1
            assert d.input(met[0, t]) == 0
        # callback returns 0: the callee owns the malloc'd buffer and must free it (:97-104)
        taken = []
        CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint8, ctypes.c_void_p)

        def cb(data, ln, ud):
            taken.append((data, bytes((ctypes.c_uint8 * code.out_bytes).from_address(data))))
            return 0
        keep = CB(cb)
        d.lib.decoder_register_callback(d.obj, ctypes.cast(keep, ctypes.c_void_p))
        for f in range(2):
            for t in range(code.T):
                assert d.input(met[f, t]) == 0
        assert len(taken) == 2 and taken[0][0] != taken[1][0]
        for f, (ptr, data) in enumerate(taken):
            assert np.array_equal(np.frombuffer(data, np.uint8), exp[f])
            libc.free(ptr)
        d.destroy()


@pytest.mark.parametrize("idx", [0, 1, 5])
def test_bsc_whole_frame_calls(cc, orc, idx):
    """BSC driver pattern: T symbol bytes in one call, callback carries the path metric (main.c:70,74)"""
    code = cc.get_code(idx)
    _, sym = make_hard_inputs(orc, ocode(code), 8, 0.06, seed=idx)
    exp, exp_pm = orc.decode_hard(ocode(code), sym)
    for lib in _targets("bsc"):
        d = cc.LegacyDecoder("bsc", lib)
        assert d.init(code) == 0
        d.register_callback()
        for f in range(8):
            if f % 2:
                assert d.input(sym[f]) == 0
            else:                                                     # split delivery is allowed too
                assert d.input(sym[f, :17]) == 0 and d.input(sym[f, 17:]) == 0
        got = np.frombuffer(b"".join(fr[0] for fr in d.frames), np.uint8).reshape(8, code.out_bytes)
        assert np.array_equal(got, exp)
        assert [fr[2] for fr in d.frames] == exp_pm.tolist() and all(fr[1] == code.L for fr in d.frames)
        d.destroy()


def test_k7_through_the_legacy_abi_at_its_width_limit(cc, orc):
    """L + K - 1 = 255 is the most the 8-bit legacy ABI can carry (code.h:12, decoder.h:26)"""
    code = cc.codes.K7_R12.with_block_len(249)
    _, met = make_soft_inputs(orc, ocode(code), 2, 0.45, seed=5)
    exp = orc.decode_soft(ocode(code), met)
    for lib in _targets("awgn"):
        d = cc.LegacyDecoder("awgn", lib)
        assert d.init(code) == 0
        d.register_callback()
        for f in range(2):
            for t in range(code.T):
                d.input(met[f, t])
        got = np.frombuffer(b"".join(fr[0] for fr in d.frames), np.uint8).reshape(2, code.out_bytes)
        assert np.array_equal(got, exp)
        d.destroy()


# ---- the reference's UNMODIFIED drivers linked against the shims (oracle/Makefile `dropin`) ----------
PUBLISHED_AWGN_SOFT = {   # results/awgn_channel.m:15,24,27 — Eb/N0 = 0, 2, 4 dB (8e5 bits per point)
    0: [0.35548875, 0.20219750, 0.05668375],
    3: [0.16392500, 0.04393375, 0.00444625],
    5: [0.11558875, 0.03868500, 0.00832625],
}


@pytest.mark.parametrize("code_idx", [0, 3, 5])
def test_unmodified_awgn_driver_on_gpu_reproduces_published_ber(code_idx):
    """AWGN-channel/main.c, byte-for-byte, with libvdec_awgn.so in the decoder slot: the BER it prints for
    the first operating points must agree with the author's published table (unseeded Monte-Carlo on both
    sides, 8e5 bits per point => a few % of statistical spread; errors are bursty, so the band is generous)."""
    import re
    import subprocess
    exe = os.path.join(REFDIR, "awgn-simulation-b200")
    if not os.path.exists(exe):
        pytest.skip("drop-in driver not built (needs /root/reference at build time)")
    # srand(time(NULL)) of the driver pinned through an LD_PRELOAD time(): the run is reproducible (oracle/fixed_time.c)
    env = dict(os.environ, LD_PRELOAD=os.path.join(REFDIR, "libfixedtime.so"), VDEC_FIXED_TIME="20260101")
    assert os.path.exists(env["LD_PRELOAD"])
    try:
        out = subprocess.run([exe, str(code_idx)], capture_output=True, text=True, timeout=10, env=env).stdout
    except subprocess.TimeoutExpired as e:      # the full 8-point sweep takes minutes; three points are enough
        out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
    assert "Error:" not in out, out[-400:]
    ber = [float(x) for x in re.findall(r"Overall transmission error rate: ([0-9.]+)", out)]
    assert len(ber) >= 3, out[-400:]
    for got, pub in zip(ber[:3], PUBLISHED_AWGN_SOFT[code_idx]):
        # The published side is an unseeded Monte-Carlo estimate from 8e5 bits with bursty errors (the catastrophic
        # codebook codes most of all): at BER < 1e-2 that is only a few hundred error events, and the observed run-to-run
        # spread of the UNMODIFIED reference is +-20 % there.  Our side is pinned (fixed rand() seed, bit-exact decoder), so
        # the comparison is deterministic.  Bit-exactness is established by the seeded tests; this one shows that the
        # unmodified driver works with the GPU decoder in its decoder slot.
        tol = 0.12 * pub + 3e-4 if pub >= 2e-2 else 0.40 * pub + 5e-4
        assert abs(got - pub) <= tol, (ber[:3], PUBLISHED_AWGN_SOFT[code_idx])


def test_unmodified_bsc_driver_runs_on_gpu():
    """binary-symmetric-channel/main.c with libvdec_bsc.so: a full point is 2e7 frames (minutes through a
    one-frame-per-call ABI), so only check that it runs without tripping the driver's self-checks."""
    import subprocess
    exe = os.path.join(REFDIR, "binary-symmetric-simulation-b200")
    if not os.path.exists(exe):
        pytest.skip("drop-in driver not built (needs /root/reference at build time)")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=6).stdout
    except subprocess.TimeoutExpired as e:
        out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
    assert "Channel error rate: 0.000001" in out and "Error:" not in out, out[-300:]


def test_c_sweep_driver_binary(cc, tmp_path):
    """examples/vdec_sweep_main.c (built as lib/vdec-sweep): the reference drivers' operating-point loop in plain C on
    top of vdec_sweep().  Its printed counters equal the library call's, and the Octave vectors carry the GPU count."""
    import re
    import subprocess
    exe = os.path.join(ROOT, "convolutional-codes_b200", "lib", "vdec-sweep")
    assert os.path.exists(exe), "run `python __graft_entry__.py build`"
    out_m = str(tmp_path / "sweep.m")
    r = subprocess.run([exe, "--code", "cb3", "--ebn0", "2:6:2", "--frames", "60000", "--gpus", "1", "--seed", "99", "--out", out_m],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-400:]
    got = [[int(x) for x in m] for m in re.findall(r"bits (\d+)  bit errors (\d+)  frame errors (\d+)", r.stdout)]
    exp = cc.sweep(cc.get_code(3), [2.0, 4.0, 6.0], 60000, cc.CHANNEL_AWGN, cc.PARITY_REF, seed=99, ndev=1)
    assert got == exp.tolist(), (got, exp.tolist())
    assert len(re.findall(r"Overall transmission error rate: [0-9.]+", r.stdout)) == 3          # the drivers' wording, main.c:233
    text = open(out_m).read()
    assert "n_gpus_cb3_ref_soft = 1;" in text and "ber_cb3_ref_soft = [" in text and "snr_cb3_ref_soft = [2 4 6];" in text
    # the 17 channel error rates of the BSC driver
    r = subprocess.run([exe, "--code", "cb0", "--channel", "bsc", "--frames", "20000", "--gpus", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.count("Channel error rate:") == 17, r.stdout[-300:]
