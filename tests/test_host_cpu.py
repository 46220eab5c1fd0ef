"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/viterbi_b200.h declares, argument
checking works without a GPU, nothing routes through a CPU fallback, and the multi-process sharding helpers behave (gloo, world 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from viterbidecodercpp_b200 import _lib, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUPPORTED_K = {3, 5, 7, 9, 15}    # constraint lengths with compiled kernels so far (grows as kernel families land)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "viterbi_b200.h")).read()
    declared = set(re.findall(r"\b(vitb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"vitb_status", "vitb_params", "vitb_decoder", "vitb_batch_opts"}
    assert len(declared) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/viterbi_b200.h but not exported"
    bound = {n for n, _, _ in _lib.API}
    assert declared == bound, f"python binding out of sync: {declared ^ bound}"


def test_headers_compile_as_c99_and_cpp17(tmp_path):
    """the drop-in boundary is a C ABI: include/viterbi_b200.h must be plain C; the facade on top of it is header-only C++17"""
    inc = os.path.join(ROOT, "include")
    c = tmp_path / "t.c"
    c.write_text('#include "viterbi_b200.h"\nint main(void) { return VITB_MAX_R > 0 && VITB_END_STATE_BEST != 0 ? 0 : 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(c)], check=True)
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "viterbi_cuda/viterbi_decoder_cuda.h"\nint main() { return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-fsyntax-only", "-I", inc, str(cpp)], check=True)


def test_library_is_built_for_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_is_supported_matches_catalogue():
    lib = v.load_library()
    for code in v.COMMON_CODES:
        for name, factory in v.DECODE_TYPES.items():
            dc = factory(code.R)
            bt = v.ViterbiBranchTable(code.K, code.R, code.G, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes)
            assert v.ViterbiDecoder_CUDA.is_valid(bt, dc.decoder_config) == (code.K in SUPPORTED_K), (code.name, name)
    # polynomials outside the compiled catalogue: the generic kernels cover 3 <= K <= 7 with R <= 6, nothing beyond
    bt = v.ViterbiBranchTable(7, 2, [0o133, 0o165], 127, -127)
    assert v.ViterbiDecoder_CUDA.is_valid(bt, v.get_soft16_decoding_config(2).decoder_config)
    bt = v.ViterbiBranchTable(9, 2, [0o561, 0o753 ^ 2], 127, -127)
    assert not v.ViterbiDecoder_CUDA.is_valid(bt, v.get_soft16_decoding_config(2).decoder_config)
    bt = v.ViterbiBranchTable(5, 7, [19, 21, 23, 25, 27, 29, 31], 127, -127)
    assert not v.ViterbiDecoder_CUDA.is_valid(bt, v.get_soft16_decoding_config(7).decoder_config)
    assert lib.vitb_version().startswith(b"viterbi_b200")


@pytest.mark.skipif(v.load_library().vitb_device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_gpu():
    dc = v.get_soft16_decoding_config(2)
    bt = v.ViterbiBranchTable(7, 2, [109, 79], 127, -127)
    with pytest.raises(v.ViterbiError) as e:
        v.ViterbiDecoder_CUDA(bt, dc.decoder_config)
    assert e.value.status == _lib.VITB_ERR_CUDA


def test_create_argument_checking():
    lib = v.load_library()
    p = _lib.vitb_params()
    h = C.c_void_p()
    p.K, p.R, p.soft_bytes, p.soft_decision_high, p.soft_decision_low = 7, 2, 3, 127, -127
    assert lib.vitb_create(C.byref(p), C.byref(h)) == _lib.VITB_ERR_ARG          # soft_bytes must be 1 or 2
    p.soft_bytes, p.soft_decision_high, p.soft_decision_low = 2, -5, 5
    assert lib.vitb_create(C.byref(p), C.byref(h)) == _lib.VITB_ERR_ARG          # high must exceed low (viterbi_branch_table.h:43)
    p.soft_decision_high, p.soft_decision_low = 127, -127
    p.K, p.G[0], p.G[1] = 9, 1, 3
    assert lib.vitb_create(C.byref(p), C.byref(h)) == _lib.VITB_ERR_UNSUPPORTED  # uncatalogued polynomials beyond the generic kernels (K <= 7)
    p.K = 7
    assert lib.vitb_create(None, C.byref(h)) == _lib.VITB_ERR_ARG
    assert lib.vitb_status_string(_lib.VITB_ERR_UNSUPPORTED)


def test_branch_table_mirror_matches_reference_definition():
    """BT[i][j] = parity((j << 1) & G[i]) ? high : low  (viterbi_branch_table.h:45-54), checked against the oracle's table"""
    import oracle_binding as ob
    for code in v.COMMON_CODES:
        bt = v.ViterbiBranchTable(code.K, code.R, code.G, 127, -127)
        ora = ob.OracleDecoder(code.K, code.R, code.G, 2, 127, -127, [0, 0, 0, 0])
        ref = np.zeros((code.R, bt.NUMSTATES), dtype=np.int32)
        ora.L.vo_branch_table(ora.h, ref.ctypes.data)
        for i in range(code.R):
            assert (bt[i] == ref[i]).all()


def test_presets_match_reference_values():
    """SOFT16 R=2: 508/0/2540/62995, R=4: 1016/0/5080/60455, R=6: 1524/0/7620/57915; HARD8 R=2: 4/0/12/243"""
    def tup(dc):
        c = dc.decoder_config
        return (c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold)
    assert tup(v.get_soft16_decoding_config(2)) == (508, 0, 2540, 62995)
    assert tup(v.get_soft16_decoding_config(4)) == (1016, 0, 5080, 60455)
    assert tup(v.get_soft16_decoding_config(6)) == (1524, 0, 7620, 57915)
    assert tup(v.get_hard8_decoding_config(2)) == (4, 0, 12, 243)
    assert tup(v.get_soft8_decoding_config(2)) == (12, 0, 24, 231)


def test_exchange_swizzle_is_conflict_free():
    """the shared-memory exchange of csrc/acs_group.cuh (GroupShape::slot): every warp-wide store and load hits 32 distinct banks
    and the slots form a bijection, for every (state bits, lanes per pair) the kernels are compiled for"""
    def check(SB, g):
        T, LB = 1 << g, SB - g
        NL = 1 << LB

        def swz(qp):
            return ((qp >> (LB - g)) & (T - 1)) if LB >= g else (qp & 31)

        def slot(p, phi):
            qp, tp = phi >> g, phi & (T - 1)
            return qp * 32 + ((p * T + tp) ^ swz(qp))
        seen = set()
        for q in range(NL):
            wr = {slot(lane // T, ((lane % T) << LB) | q) % 32 for lane in range(32)}
            rd = {slot(lane // T, (q << g) | (lane % T)) % 32 for lane in range(32)}
            assert len(wr) == 32 and len(rd) == 32, (SB, g, q)
            seen |= {slot(lane // T, (q << g) | (lane % T)) for lane in range(32)}
        assert seen == set(range(NL * 32))
    for SB, g in [(4, 2), (6, 1), (6, 2), (6, 3), (8, 3), (8, 4), (8, 5)]:
        check(SB, g)


def test_hist_group_exchange_layout():
    """csrc/acs_hist_group.cuh (HistGroupShape::slot), one frame over 4 lanes, K = 9 (64 registers per lane) and K = 7 (16): the slots
    are a bijection; a lane's four consecutive registers are one aligned 16-byte group and the 8 lanes of a quarter warp write 8
    different groups (STS.128 is served a quarter warp at a time); every warp-wide load hits 32 distinct banks; and the base addresses
    the kernel uses - one per lane for the stores, four for the loads, plus compile-time offsets - are that slot function"""
    g = 2
    T = 1 << g
    for LB in (6, 4):
        NL = 1 << LB

        def slot(fw, phi):
            qp, tp = phi >> g, phi & (T - 1)
            return qp * 32 + ((fw + 2 * (qp >> (LB - g))) & 7) * 4 + tp
        assert {slot(lane >> g, (q << g) | (lane & 3)) for lane in range(32) for q in range(NL)} == set(range(NL * 32))
        for q in range(NL):
            assert len({slot(lane >> g, (q << g) | (lane & 3)) % 32 for lane in range(32)}) == 32
        for q4 in range(0, NL, 4):
            for quarter in range(4):
                words = [slot(lane >> g, ((lane & 3) << LB) | q4) for lane in range(8 * quarter, 8 * quarter + 8)]
                assert all(w % 4 == 0 for w in words) and len({(w % 32) // 4 for w in words}) == 8
        for lane in range(32):
            fw, t = lane >> g, lane & (T - 1)
            wr_base = (t << (LB - g)) * 32 + ((fw + 2 * t) & 7) * 4
            rd_base = [((fw + 2 * k) & 7) * 4 + t for k in range(4)]
            for q in range(NL):
                assert wr_base + (q >> 2) * 32 + (q & 3) == slot(fw, (t << LB) | q)
                assert rd_base[(q >> (LB - g)) & 3] + q * 32 == slot(fw, (q << g) | t)


def test_hist_group_record_position_of_a_state():
    """position algebra shared by acs_hist_group_kernel and traceback_hist_kernel: after n steps since the last exchange logical state s
    sits at PHI = rotr^n(s) (8-bit rotation); an in-place butterfly at phase n pairs positions that differ in bit 7 - n and leaves new
    state 2j at the position of old j, 2j + 1 at the position of old j + 128"""
    SB = 8
    mask = (1 << SB) - 1
    rotr = lambda v, r: ((v >> (r % SB)) | (v << (SB - r % SB))) & mask if r % SB else v
    for n in range(6):
        for j in range(128):
            p0, p1 = rotr(j, n), rotr(j + 128, n)
            assert p0 ^ p1 == 1 << (SB - 1 - n)                      # partners differ in the phase bit
            assert rotr((2 * j) & mask, n + 1) == p0                 # new state 2j stays where old j was
            assert rotr((2 * j + 1) & mask, n + 1) == p1             # new state 2j+1 goes where old j+128 was


def test_cta_exchange_skew_is_conflict_free():
    """csrc/acs_cta.cuh CtaShape::slot(phi) = phi + 4 * (phi >> 5): the exchange's loads ((q << LOGT) | t, fixed q) hit 32 distinct
    banks per warp, its stores - 2^LB consecutive words per thread, as STS.128, served a quarter warp at a time - 8 distinct 16-byte
    bank groups per quarter warp (and 32 distinct banks per warp as scalar stores), slots never collide and stay 16-byte aligned"""
    SB = 14
    for LOGT in (9, 10):
        LB, T = SB - LOGT, 1 << LOGT
        NL = 1 << LB
        slot = lambda phi: phi + ((phi >> 5) << 2)
        assert len({slot(p) for p in range(1 << SB)}) == 1 << SB
        assert max(slot(p) for p in range(1 << SB)) < (1 << SB) + (1 << SB) // 8 + 32
        for w in range(T // 32):
            lanes = range(32 * w, 32 * w + 32)
            for q in range(NL):
                assert len({slot((q << LOGT) | t) % 32 for t in lanes}) == 32
            for q4 in range(0, NL, 4):
                for quarter in range(4):
                    ts = range(32 * w + 8 * quarter, 32 * w + 8 * quarter + 8)
                    words = [slot((t << LB) | q4) for t in ts]
                    assert all(x % 4 == 0 for x in words)
                    assert len({(x % 32) // 4 for x in words}) == 8
        # the split form the kernels use (per-thread offset + compile-time offset: CtaShape::rd_off / wr_off / RD_STRIDE)
        for t in range(T):
            for q in range(NL):
                assert slot((q << LOGT) | t) == (t + ((t >> 5) << 2)) + q * (T + T // 8)
                assert slot((t << LB) | q) == slot(t << LB) + q


def test_frame_range_partitions_exactly():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 64, 65536, 1000003):
            spans = [sharding.frame_range(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from viterbidecodercpp_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
f0, f1 = sharding.frame_range(rank, world, 1001)
total = sharding.sum_over_ranks(f1 - f0)
worst = sharding.max_over_ranks(10.0 + rank)
dist.barrier()
assert total == 1001 and worst == 10.0 + world - 1, (total, worst)
if rank == 0:
    print("OK", int(total), worst)
dist.destroy_process_group()
"""


def test_two_rank_gloo_sharding(tmp_path):
    """the N>1 plumbing bench.py uses (shard frames, barrier, max over ranks) with world_size 2 on CPU"""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script), ROOT], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "OK 1001 11.0" in r.stdout


def test_tie_break_flavours_are_checked_without_a_gpu():
    """VITB_TIE_SIMD_SAT exists for the catalogue codes only (decision-row kernels); unknown flavours are rejected"""
    lib = v.load_library()
    p = _lib.vitb_params()
    p.K, p.R, p.soft_bytes, p.soft_decision_high, p.soft_decision_low = 7, 2, 1, 3, -3
    p.G[0], p.G[1] = 109, 79
    p.soft_decision_max_error, p.initial_non_start_error, p.renormalisation_threshold = 12, 36, 200
    for tie, want in ((_lib.VITB_TIE_SCALAR, 1), (_lib.VITB_TIE_SIMD, 1), (_lib.VITB_TIE_SIMD_SAT, 1), (3, 0), (-1, 0)):
        p.tie_break = tie
        assert lib.vitb_is_supported(C.byref(p)) == want, tie
    p.G[1] = 0o165                       # uncatalogued polynomials: generic kernel, which has no saturating flavour
    for tie, want in ((_lib.VITB_TIE_SCALAR, 1), (_lib.VITB_TIE_SIMD, 1), (_lib.VITB_TIE_SIMD_SAT, 0)):
        p.tie_break = tie
        assert lib.vitb_is_supported(C.byref(p)) == want, tie


@pytest.mark.skipif(not os.path.isdir("/root/reference/include/viterbi"), reason="needs the reference headers (build container only)")
def test_reference_adapter_and_cuda_slot_compile_against_the_reference_headers(tmp_path):
    """include/viterbi_cuda/viterbi_decoder_cuda_ref.h (decoder class on the reference's own Core) and the test-only SIMD_CUDA slot
    (oracle/ref_programs/cuda_slot.h) against the unmodified reference tree; the facade's converting constructors too"""
    inc = os.path.join(ROOT, "include")
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include <stddef.h>\n#include <stdint.h>\n#include "cuda_slot.h"\n#include "viterbi_cuda/viterbi_decoder_cuda.h"\n'
                   'using D = viterbi_cuda::ViterbiDecoder_CUDA_Ref<7, 2, uint16_t, int16_t>;\n'
                   'static_assert(D::is_valid, "u16/s16 is a supported type pair");\n'
                   'static_assert(!viterbi_cuda::ViterbiDecoder_CUDA_Ref<7, 2, uint32_t, int16_t>::is_valid, "other pairs are not");\n'
                   'uint64_t f(D::Base& b, const int16_t* s, size_t n) { return D::update<uint64_t>(b, s, n); }\n'
                   'viterbi_cuda::ViterbiBranchTable<7, 2, int16_t> g(const ::ViterbiBranchTable<7, 2, int16_t>& t) { return viterbi_cuda::ViterbiBranchTable<7, 2, int16_t>(t); }\n'
                   'int main() { return simd_type_list_with_cuda().back() == SIMD_CUDA ? 0 : 1; }\n')
    subprocess.run(["g++", "-std=c++17", "-mavx2", "-msse4.2", "-Wall", "-fsyntax-only", "-I", inc, "-I", "/root/reference/include", "-I", "/root/reference/examples",
                    "-I", os.path.join(ROOT, "oracle", "ref_programs"), str(cpp)], check=True)


def test_k15_table_slot_maps_compile_time_checks(tmp_path):
    """tests/host_programs/pair_map_check.cu: the slot maps of the K = 15 kernels' table fetch (csrc/acs_cta.cuh PairMap) are
    bijections, put the butterflies of a pair / quad into one 16-byte slot and keep quarter warps free of bank conflicts - checked by
    static_asserts over the same constexpr evaluation the kernels use, so compiling the program IS the test"""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = tmp_path / "pair_map_check"
    subprocess.check_call([nvcc, "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                           "-I", os.path.join(ROOT, "viterbidecodercpp_b200", "csrc"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "host_programs", "pair_map_check.cu")])
    assert "pair maps ok" in subprocess.check_output([str(exe)]).decode()
