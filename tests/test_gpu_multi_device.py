"""vitb_decode_batch_multi on REAL devices (one handle per GPU, contiguous frame ranges, host scatter, no
collective): equality with the single-device result and with the oracle, from pageable memory (one host thread per device) and
from pinned memory (one host thread enqueues every device and joins).  Skipped below two visible devices; the single-device
variant of the test (two handles on device 0) is tests/test_gpu_parity.py::test_decode_batch_multi_matches_single."""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from common import CODE_BY_NAME, assert_batch_equal, frames, make_cuda_decoder, make_oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def n_devices():
    return v.load_library().vitb_device_count()


@pytest.mark.parametrize("name,decode_type,n_frames,L", [("Voyager", "HARD8", 9001, 512), ("CDMA IS-95A", "SOFT16", 1111, 256),
                                                           ("Cassini", "SOFT16", 9, 256)])
def test_multi_device_equals_single_device_and_oracle(cuda_lib, name, decode_type, n_frames, L):
    n = n_devices()
    if n < 2:
        pytest.skip("needs at least two CUDA devices")
    code = CODE_BY_NAME[name]
    decs = [make_cuda_decoder(code, decode_type, device=d)[0] for d in range(n)]
    dc = v.DECODE_TYPES[decode_type](code.R)
    ora, _ = make_oracle(code, decode_type)
    tx, sym = frames(code, dc, n_frames, L, 2.0, seed=77)
    want = ora.decode_frames(sym, n_frames, L)
    single = decs[0].decode_batch(sym, L)
    assert_batch_equal(single, want, f"{name} single device")
    got = v.decode_batch_multi(decs, sym, L)                       # pageable numpy memory: one host thread per device
    assert_batch_equal(got, want, f"{name} over {n} devices, pageable host memory")
    # pinned memory: single-threaded asynchronous enqueue on every device
    h_sym = torch.from_numpy(sym).pin_memory()
    h_out = torch.zeros((n_frames, (L + 7) // 8), dtype=torch.uint8).pin_memory()
    h_acc = torch.zeros(n_frames, dtype=torch.int64).pin_memory()
    h_fin = torch.zeros(n_frames, dtype=torch.int32).pin_memory()
    for _ in range(2):                                             # twice: the second call reuses every workspace
        h_out.zero_()
        v.decode_batch_multi_raw(decs, h_sym.data_ptr(), n_frames, L, h_out.data_ptr(), h_acc.data_ptr(), h_fin.data_ptr(),
                                 row_stride=sym.shape[1])
        got = (h_out.numpy(), h_acc.numpy().astype(np.uint64), h_fin.numpy().astype(np.uint32))
        assert_batch_equal(got, want, f"{name} over {n} devices, pinned host memory")
    for d in decs:
        d.close()


def test_more_handles_than_frames(cuda_lib):
    """ranges may be empty: 3 frames over every device (and over 5 handles on one device)"""
    n = max(n_devices(), 1)
    code = CODE_BY_NAME["Voyager"]
    decs = [make_cuda_decoder(code, "SOFT16", device=d % n)[0] for d in range(max(n, 5))]
    dc = v.DECODE_TYPES["SOFT16"](code.R)
    ora, _ = make_oracle(code, "SOFT16")
    tx, sym = frames(code, dc, 3, 128, 3.0, seed=3)
    want = ora.decode_frames(sym, 3, 128)
    assert_batch_equal(v.decode_batch_multi(decs, sym, 128), want, "3 frames over 5+ handles")
    for d in decs:
        d.close()
