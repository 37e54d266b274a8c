"""CPU: host-side logic, the C ABI surface and the 2-rank counter exchange (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, ocode

ALL_CODES = lambda cc: [cc.get_code(i) for i in (0, 1, 2, 3, 5)] + [cc.codes.K7_R12, cc.codes.K7_R13, cc.codes.K8_R12, cc.codes.K9_R12, cc.codes.K10_R12, cc.codes.K11_R12, cc.codes.K15_R12]


def test_labels_match_oracle(cc, orc):
    for code in ALL_CODES(cc):
        for parity in (cc.PARITY_REF, cc.PARITY_LINEAR):
            assert np.array_equal(cc.config_labels(code, parity), orc.labels(ocode(code, parity))), (code.name, parity)


def test_ref_labels_are_nonlinear_only_with_a_d3_tap(cc):
    # SURVEY finding 1: codes 0 and 5 (K=3) are unaffected, (171,133) is
    for idx in (0, 5):
        c = cc.get_code(idx)
        assert np.array_equal(cc.config_labels(c, 0), cc.config_labels(c, 1))
    c = cc.codes.K7_R12
    assert not np.array_equal(cc.config_labels(c, 0), cc.config_labels(c, 1))


def test_kernel_selection(cc):
    k7 = cc.codes.K7_R12
    assert cc.config_kernel_name(k7, cc.SOFT_F32, cc.PARITY_REF).startswith("regs_soft_f32<K7 (171,133) ref")
    assert cc.config_kernel_name(k7, cc.SOFT_F32, cc.PARITY_LINEAR).startswith("regs_soft_f32<K7 (171,133) linear")
    assert cc.config_kernel_name(k7, cc.SOFT_F32, cc.PARITY_REF, cc.FORCE_GENERIC).startswith("generic_soft_f32")
    # K = 8 .. 12: lane-group register kernel for linear labels, table-driven for the reference's quirk labels
    assert cc.config_kernel_name(cc.codes.K9_R12, cc.SOFT_F32, cc.PARITY_REF).startswith("tile_soft_f32<K9 (561,753) ref")
    # no built-in instantiation: compiled at run time for the polynomials (NVRTC), table-driven only on request
    assert cc.config_kernel_name(cc.codes.K10_R12, cc.SOFT_F32, cc.PARITY_REF) == "tile_soft_f32<K10 (1167,1545) ref-labels, nvrtc>"
    assert cc.config_kernel_name(cc.codes.K10_R12, cc.SOFT_F32, cc.PARITY_REF, cc.NO_JIT).startswith("generic_soft_f32<K10")
    assert cc.config_kernel_name(cc.codes.K8_R12, cc.SOFT_F32, cc.PARITY_LINEAR).startswith("tile_soft_f32<K8")
    assert cc.config_kernel_name(cc.codes.K8_R12, cc.SOFT_F32, cc.PARITY_REF).startswith("tile_soft_f32<K8 (371,247) ref")
    assert cc.config_kernel_name(cc.codes.K8_R12, cc.SOFT_F32, cc.PARITY_LINEAR, cc.PREFER_LANEGROUP).startswith("regsw_soft_f32<K8")
    assert cc.config_kernel_name(cc.codes.K9_R12, cc.SOFT_F32, cc.PARITY_LINEAR).startswith("tile_soft_f32<K9")
    assert cc.config_kernel_name(cc.codes.K9_R12, cc.SOFT_F32, cc.PARITY_LINEAR, cc.PREFER_LANEGROUP).startswith("regsw_soft_f32<K9")
    assert cc.config_kernel_name(cc.codes.K11_R12, cc.SOFT_F32, cc.PARITY_LINEAR, cc.PREFER_LANEGROUP).startswith("regsw_soft_f32<K11")
    assert cc.config_kernel_name(cc.codes.K13_R12, cc.SOFT_F32, cc.PARITY_LINEAR).startswith("tile_soft_f32<K13")
    assert cc.config_kernel_name(cc.codes.K15_R12, cc.SOFT_F32, cc.PARITY_LINEAR).startswith("tile_soft_f32<K15")
    assert cc.config_kernel_name(cc.codes.K15_R12, cc.SOFT_F32, cc.PARITY_REF).startswith("tile_soft_f32<K15 (46321,51271) ref")
    assert cc.config_kernel_name(cc.codes.K15_R12, cc.SOFT_F32, cc.PARITY_REF, cc.FORCE_GENERIC).startswith("generic_soft_f32<K15")
    # hard decisions: 32-bit register kernel; packed 16-bit kernel only for labellings the former lacks
    assert cc.config_kernel_name(cc.get_code(0), cc.HARD_U16).startswith("regs2_hard_u16x2<codebook0")
    assert cc.config_kernel_name(cc.get_code(0), cc.HARD_U16, cc.PARITY_REF, cc.PREFER_32BIT).startswith("regs_hard_u16<codebook0")
    assert cc.config_kernel_name(cc.codes.K7_R13, cc.HARD_U16).startswith("regs2_hard_u16x2<K7 (133,171,165) ref")
    # frames too long for the shared-memory output tile go to the table-driven kernels
    assert cc.config_kernel_name(cc.get_code(0).with_block_len(20000), cc.HARD_U16).startswith("generic_hard_u16")
    # integer metrics on large trellises: packed 16-bit register-tile kernels
    assert cc.config_kernel_name(cc.codes.K8_R12, cc.HARD_U16, cc.PARITY_REF).startswith("tile2_hard_u16x2<K8 (371,247) ref")
    assert cc.config_kernel_name(cc.codes.K9_R12, cc.HARD_U16, cc.PARITY_LINEAR).startswith("tile2_hard_u16x2<K9")
    assert cc.config_kernel_name(cc.codes.K15_R12, cc.HARD_U16, cc.PARITY_LINEAR).startswith("tile2_hard_u16x2<K15")
    assert cc.config_kernel_name(cc.codes.K15_R12, cc.HARD_U16, cc.PARITY_REF) == "tile2_hard_u16x2<K15 (46321,51271) ref-labels, nvrtc>"
    assert cc.config_kernel_name(cc.codes.K15_R12, cc.HARD_U16, cc.PARITY_REF, cc.NO_JIT).startswith("generic_hard_u16<K15")
    assert cc.config_kernel_name(cc.codes.K9_R12, cc.SOFT_U8, cc.PARITY_LINEAR, metric_max=7).startswith("tile2_soft_u8x2<K9")
    assert cc.config_kernel_name(cc.codes.K9_R12, cc.SOFT_U8, cc.PARITY_LINEAR, metric_max=255).startswith("generic_soft_u8<K9")
    # quantised soft mode: packed kernel if the caller bounds the metrics, table-driven otherwise
    assert cc.config_kernel_name(k7, cc.SOFT_U8, metric_max=2).startswith("regs2_soft_u8x2<K7 (171,133) ref")
    assert cc.config_kernel_name(k7, cc.SOFT_U8).startswith("generic_soft_u8")
    # codebook 0 has no D^3 tap: the ref-label kernel also serves the linear request
    assert cc.config_kernel_name(cc.get_code(0), cc.SOFT_F32, cc.PARITY_LINEAR).startswith("regs_soft_f32<codebook0")


def test_unregistered_polynomials_compile_at_run_time(cc):
    """NVRTC needs no GPU: the four kernel families instantiate for polynomials outside the registries"""
    from convolutional_codes_b200.codes import CodeParam, left_align
    k7 = CodeParam(2, 7, 300, (left_align("165", 7), left_align("127", 7)))
    k6r3 = CodeParam(3, 6, 100, (left_align("47", 6), left_align("53", 6), left_align("75", 6)))
    k9 = CodeParam(2, 9, 200, (left_align("657", 9), left_align("435", 9)))
    assert cc.config_kernel_name(k7, cc.SOFT_F32, cc.PARITY_LINEAR) == "regs_soft_f32<K7 (165,127) linear, nvrtc>"
    assert cc.config_kernel_name(k7, cc.HARD_U16, cc.PARITY_REF) == "regs2_hard_u16x2<K7 (165,127) ref-labels, nvrtc>"
    assert cc.config_kernel_name(k7, cc.HARD_U16, cc.PARITY_REF, cc.PREFER_32BIT) == "regs_hard_u16<K7 (165,127) ref-labels, nvrtc>"
    assert cc.config_kernel_name(k9, cc.SOFT_U8, cc.PARITY_LINEAR, metric_max=3) == "tile2_soft_u8x2<K9 (657,435) linear, nvrtc>"
    assert cc.config_kernel_name(k9, cc.SOFT_F32, cc.PARITY_LINEAR, cc.FORCE_GENERIC).startswith("generic_soft_f32")
    assert cc.jit_check(cc.codes.K7_R12, cc.SOFT_F32, cc.PARITY_REF) == 0           # built in: nothing to compile
    assert cc.jit_check(k7, cc.SOFT_F32, cc.PARITY_LINEAR, cc.NO_JIT) == 0
    assert cc.jit_check(k7, cc.SOFT_F32, cc.PARITY_LINEAR) > 10000                  # a cubin came back
    assert cc.jit_check(k6r3, cc.HARD_U16, cc.PARITY_LINEAR) > 10000
    assert cc.jit_check(k7, cc.SOFT_F32, cc.PARITY_LINEAR, iq=True) > 10000         # raw-I/Q variant: demapper in the feeder
    assert cc.jit_check(k6r3, cc.SOFT_F32, cc.PARITY_REF, iq=True) > 10000
    assert cc.jit_check(k9, cc.SOFT_F32, cc.PARITY_LINEAR, iq=True) == 0            # tile kernels take staged metrics
    assert cc.jit_check(k9, cc.SOFT_F32, cc.PARITY_REF) > 10000
    # rate 1/3 on a large trellis: register-tile kernel with an 8-letter label alphabet (true-parity labels), built in
    # for (557,663,711) and compiled at run time for other polynomials; de-facto labels with a D^3-type tap and integer
    # metrics stay table-driven
    k9r3 = CodeParam(3, 9, 100, (left_align("557", 9), left_align("663", 9), left_align("711", 9)))
    assert cc.config_kernel_name(k9r3, cc.SOFT_F32, cc.PARITY_LINEAR).startswith("tile_soft_f32<K9 r1/3 (557,663,711)")
    k10r3 = CodeParam(3, 10, 100, (left_align("1117", 10), left_align("1365", 10), left_align("1633", 10)))
    assert cc.config_kernel_name(k10r3, cc.SOFT_F32, cc.PARITY_LINEAR) == "tile_soft_f32<K10 (1117,1365,1633) linear, nvrtc>"
    assert cc.jit_check(k10r3, cc.SOFT_F32, cc.PARITY_LINEAR) > 10000
    assert cc.config_kernel_name(k9r3, cc.HARD_U16, cc.PARITY_LINEAR).startswith("generic_hard_u16<K9 n3")


def _declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vdec_[a-z_]+|decoder_[a-z_]+)\s*\(", src)))


def test_c_abi_exports_every_declared_symbol(cc):
    libdir = os.path.join(ROOT, "convolutional-codes_b200", "lib")
    lib = ctypes.CDLL(os.path.join(libdir, "libvdec.so"))
    names = _declared_functions("vdec.h")
    assert len(names) >= 19
    for n in names:
        getattr(lib, n)
    for kind in ("awgn", "bsc"):
        shim = ctypes.CDLL(os.path.join(libdir, f"libvdec_{kind}.so"))
        decl = _declared_functions(f"vdec_legacy_{kind}.h")
        assert decl == sorted(["decoder_create", "decoder_destroy", "decoder_init", "decoder_register_callback",
                               "decoder_reset", "decoder_input"])
        for n in decl:
            getattr(shim, n)


def test_code_param_layout_matches_reference_abi():
    """struct code_param crosses the plugin boundary by pointer: offsets must be the reference's"""
    src = r'''
    #include <stddef.h>
    #include <stdio.h>
    #include "vdec_code_param.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(struct code_param), offsetof(struct code_param, block_len),
      offsetof(struct code_param, polynomials), offsetof(struct code_param, metric_weight),
      offsetof(struct code_param, fano_metric_weight), offsetof(struct code_param, userdata)); return 0; }'''
    exe = os.path.join(ROOT, "build", "abi_probe")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    open(exe + ".c", "w").write(src)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), exe + ".c", "-o", exe], check=True)
    got = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert got == ["48", "2", "8", "32", "36", "40"]      # x86-64 SysV layout of common/include/code.h:9-19
    ref_hdr = "/root/reference/common/include"
    if os.path.isdir(ref_hdr):                              # same probe against the real header when present
        open(exe + "_ref.c", "w").write(src.replace('"vdec_code_param.h"', '"code.h"'))
        subprocess.run(["gcc", "-I", ref_hdr, exe + "_ref.c", "-o", exe + "_ref"], check=True)
        assert subprocess.run([exe + "_ref"], capture_output=True, text=True, check=True).stdout.split() == got


def test_library_refuses_without_gpu(cc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    os.environ["VDEC_QUIET"] = "1"
    with pytest.raises(cc.VdecError, match="no CUDA device"):
        cc.BatchDecoder(cc.get_code(0))
    leg = cc.LegacyDecoder("awgn")
    assert leg.init(cc.get_code(0)) == -1                   # decoder_init reports failure, no CPU fallback
    assert leg.input(np.zeros(4, np.float32)) == -1
    with pytest.raises(cc.VdecError, match="no CUDA device"):      # the C sweep entry neither
        cc.sweep(cc.get_code(0), [2.0], 100)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "convolutional-codes_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").lower() or f in (), (f, "mentions oracle")


def test_shard_range_partitions(cc):
    for total in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [cc.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (a, n), (b, _) in zip(spans, spans[1:]):
                assert a + n == b
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VDEC_ROOT"])
import vdec_pkg
cc = vdec_pkg.load()
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
r, w = dist.get_rank(), dist.get_world_size()
first, count = cc.shard_range(1001, r, w)
# each rank "decodes" its own frames: counters = [bits, bit errors, frame errors] of its shard
c = torch.tensor([count * 40, sum(f % 7 for f in range(first, first + count)), sum(1 for f in range(first, first + count) if f % 7)], dtype=torch.int64)
cc.allreduce_counters(c)
exp = [1001 * 40, sum(f % 7 for f in range(1001)), sum(1 for f in range(1001) if f % 7)]
assert c.tolist() == exp, (c.tolist(), exp)
dist.destroy_process_group()
'''


def test_two_rank_counter_allreduce_gloo(cc):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", WORLD_SIZE="2", VDEC_ROOT=ROOT)
    procs = [subprocess.Popen([sys.executable, "-c", _WORKER], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm) runs without a GPU and prints
    one JSON line with the same metric / unit / config keys plus impl, cpu_baseline and a zero-copy e2e object."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "viterbi_decoded_gbit_per_s_k7_r12_soft" and d["unit"] == "Gbit/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["config"]["workload"].startswith("awgn-simulation")
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
