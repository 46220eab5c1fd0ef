"""Two-slot / two-stream batch calls (csrc/vitb_api.cu BatchSlot): the traceback of a chunk runs on the handle's second stream next to
the add-compare-select of the next chunk.  Checked here: results of pipelined device-pointer calls (vitb_set_pipelining) equal the
oracle's and arrive when the header says they do (after the next call on that stream, or after vitb_batch_flush), different batches
in consecutive calls do not leak into each other, multi-chunk calls alternate slots correctly, and a call on another stream is
ordered behind the previous one.  Device memory comes from torch (plumbing only)."""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from common import CODE_BY_NAME, assert_batch_equal, frames, make_cuda_decoder, make_oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class DevBatch:
    def __init__(self, n_frames, L):
        self.out = torch.zeros((n_frames, (L + 7) // 8), dtype=torch.uint8, device="cuda")
        self.acc = torch.zeros(n_frames, dtype=torch.int64, device="cuda")
        self.fin = torch.zeros(n_frames, dtype=torch.int32, device="cuda")

    def ptrs(self):
        return self.out.data_ptr(), self.acc.data_ptr(), self.fin.data_ptr()

    def host(self):
        return self.out.cpu().numpy(), self.acc.cpu().numpy().astype(np.uint64), self.fin.cpu().numpy().astype(np.uint32)


@pytest.mark.parametrize("name,decode_type,n_frames,L", [("Voyager", "HARD8", 4099, 600), ("Voyager", "SOFT16", 2051, 504),
                                                           ("CDMA IS-95A", "SOFT16", 1200, 400), ("DAB Radio", "SOFT8", 777, 304)])
def test_pipelined_calls_match_the_oracle(cuda_lib, name, decode_type, n_frames, L):
    """three different batches through consecutive pipelined calls into three output sets; after the flush all three are exact"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    dec.set_pipelining(True)
    s = torch.cuda.Stream()
    batches, wants, outs = [], [], []
    for k in range(3):
        tx, sym = frames(code, dc, n_frames, L, 1.0 + k, seed=900 + k)
        wants.append(ora.decode_frames(sym, n_frames, L))
        batches.append(to_dev(sym))
        outs.append(DevBatch(n_frames, L))
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        for k in range(3):
            dec.decode_batch_dev(batches[k].data_ptr(), n_frames, L, *outs[k].ptrs(), stream=s.cuda_stream)
        dec.batch_flush(s.cuda_stream)
    s.synchronize()
    for k in range(3):
        assert_batch_equal(outs[k].host(), wants[k], f"{name} {decode_type} pipelined call {k} ({dec.kernel_name})")
    dec.close()


def test_one_output_set_reused_by_consecutive_pipelined_calls(cuda_lib):
    """four batches into ONE output set: the tracebacks run in call order, so after the flush the set holds the last batch"""
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "HARD8")
    ora, _ = make_oracle(code, "HARD8")
    n_frames, L = 8192, 1024
    dec.set_pipelining(True)
    s = torch.cuda.Stream()
    syms, wants = [], []
    for k in range(4):
        tx, sym = frames(code, dc, n_frames, L, 2.0, seed=50 + k)
        syms.append(to_dev(sym))
        wants.append(ora.decode_frames(sym, n_frames, L))
    out = DevBatch(n_frames, L)                       # ONE output set reused by every call (tracebacks run in call order)
    snaps = []
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        for k in range(4):
            dec.decode_batch_dev(syms[k].data_ptr(), n_frames, L, *out.ptrs(), stream=s.cuda_stream)
        dec.batch_flush(s.cuda_stream)                # the last call's results need the flush
        snaps.append((out.out.clone(), out.acc.clone(), out.fin.clone()))
    s.synchronize()
    got = (snaps[0][0].cpu().numpy(), snaps[0][1].cpu().numpy().astype(np.uint64), snaps[0][2].cpu().numpy().astype(np.uint32))
    assert_batch_equal(got, wants[3], "last of four pipelined calls into one output set")
    dec.close()


def test_deferred_join_makes_previous_results_visible(cuda_lib):
    """call A into outputs A, call B into outputs B, then (no flush) a copy of outputs A on the same stream: exact"""
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "SOFT16")
    ora, _ = make_oracle(code, "SOFT16")
    n_frames, L = 6000, 800
    dec.set_pipelining(True)
    s = torch.cuda.Stream()
    txa, syma = frames(code, dc, n_frames, L, 2.0, seed=1)
    txb, symb = frames(code, dc, n_frames, L, 3.0, seed=2)
    want_a = ora.decode_frames(syma, n_frames, L)
    da, db = to_dev(syma), to_dev(symb)
    oa, obb = DevBatch(n_frames, L), DevBatch(n_frames, L)
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        dec.decode_batch_dev(da.data_ptr(), n_frames, L, *oa.ptrs(), stream=s.cuda_stream)
        dec.decode_batch_dev(db.data_ptr(), n_frames, L, *obb.ptrs(), stream=s.cuda_stream)
        snap = (oa.out.clone(), oa.acc.clone(), oa.fin.clone())
    s.synchronize()
    got = (snap[0].cpu().numpy(), snap[1].cpu().numpy().astype(np.uint64), snap[2].cpu().numpy().astype(np.uint32))
    assert_batch_equal(got, want_a, "results of call A behind call B without a flush")
    dec.batch_flush(0)
    torch.cuda.synchronize()
    dec.close()


def test_default_mode_results_are_in_stream_order_at_return(cuda_lib):
    """pipelining off (the default): a copy enqueued right behind the call sees its results; calls alternate slots"""
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "HARD8")
    ora, _ = make_oracle(code, "HARD8")
    n_frames, L = 5000, 704
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for k, s in enumerate([s1, s2, s1, s2]):          # consecutive calls of one handle on two different streams
        tx, sym = frames(code, dc, n_frames, L, 2.0, seed=70 + k)
        want = ora.decode_frames(sym, n_frames, L)
        d = to_dev(sym)
        o = DevBatch(n_frames, L)
        torch.cuda.synchronize()
        with torch.cuda.stream(s):
            dec.decode_batch_dev(d.data_ptr(), n_frames, L, *o.ptrs(), stream=s.cuda_stream)
            snap = (o.out.clone(), o.acc.clone(), o.fin.clone())
        s.synchronize()
        got = (snap[0].cpu().numpy(), snap[1].cpu().numpy().astype(np.uint64), snap[2].cpu().numpy().astype(np.uint32))
        assert_batch_equal(got, want, f"default mode, call {k}")
    dec.close()


@pytest.mark.parametrize("pipelined", [False, True])
def test_multi_chunk_calls_alternate_slots(cuda_lib, pipelined):
    """a workspace limit that cuts the batch into 5+ chunks: chunk c's traceback overlaps chunk c+1's ACS inside one call"""
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "HARD8")
    ora, _ = make_oracle(code, "HARD8")
    n_frames, L = 3000, 640
    dec.set_workspace_limit(dec.workspace_bytes(640, L) * 2)       # two slots of ~640 frames each
    dec.set_pipelining(pipelined)
    tx, sym = frames(code, dc, n_frames, L, 2.0, seed=5)
    want = ora.decode_frames(sym, n_frames, L)
    d = to_dev(sym)
    o = DevBatch(n_frames, L)
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        dec.decode_batch_dev(d.data_ptr(), n_frames, L, *o.ptrs(), stream=s.cuda_stream)
        dec.batch_flush(s.cuda_stream)
    s.synchronize()
    assert_batch_equal(o.host(), want, f"multi-chunk device call, pipelined={pipelined}")
    got = dec.decode_batch(sym, L)                                  # host-pointer path: chunks + device-to-host copies on the traceback stream
    assert_batch_equal(got, want, "multi-chunk host call")
    dec.close()
