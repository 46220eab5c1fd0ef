"""Host-side mirror of the reference's interface for the hot path, over the C ABI of include/viterbi_b200.h.

Reference shapes mirrored (same names, argument meaning, error behaviour):
  ViterbiBranchTable<K,R,soft_t>(G, high, low)                   include/viterbi/viterbi_branch_table.h:19-74
  ViterbiDecoder_Config<error_t>                                 include/viterbi/viterbi_decoder_config.h:11-18
  ViterbiDecoder_Core<K,R,error_t,soft_t>(branch_table, config)  include/viterbi/viterbi_decoder_core.h:157-243
      .set_traceback_length / .get_traceback_length / .reset / .get_error / .chainback
  Decoder::update<uint64_t>(core, symbols, N)                    include/viterbi/viterbi_decoder_scalar.h:28-55
Where the reference `assert`s, this raises ViterbiError(VITB_ERR_ARG / VITB_ERR_STATE).

All computation happens in the CUDA library; this file only marshals arguments.
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import vitb_params, vitb_batch_opts, VITB_TIE_SCALAR, VITB_TIE_SIMD
from .presets import ViterbiDecoder_Config


class ViterbiError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        msg = _lib.load().vitb_status_string(status).decode()
        super().__init__(f"{what}: {msg} ({status})" if what else f"{msg} ({status})")


def _check(rc, what=""):
    if rc != 0:
        raise ViterbiError(rc, what)


def _soft_dtype(soft_bytes):
    return np.int8 if soft_bytes == 1 else np.int16


class ViterbiBranchTable:
    """viterbi_branch_table.h:19-74.  On the GPU the table is folded into the kernels; this object carries (K, R, G, high, low)
    and can materialise the table for inspection: table[i][j] = high if parity((j << 1) & G[i]) else low."""

    def __init__(self, K, R, G, soft_decision_high, soft_decision_low, soft_bytes=2):
        assert K > 1 and R > 1 and len(G) == R                      # viterbi_branch_table.h:39-40
        assert soft_decision_high > soft_decision_low               # viterbi_branch_table.h:43
        self.K, self.R, self.G = K, R, [int(g) for g in G]
        self.soft_decision_high, self.soft_decision_low = int(soft_decision_high), int(soft_decision_low)
        self.soft_bytes = soft_bytes
        self.NUMSTATES = (1 << (K - 1)) // 2                        # viterbi_branch_table.h:26

    def __getitem__(self, index):                                   # operator[] viterbi_branch_table.h:59-62
        assert 0 <= index < self.R
        j = np.arange(self.NUMSTATES, dtype=np.uint64)
        v = (j << np.uint64(1)) & np.uint64(self.G[index])
        par = np.zeros_like(v)
        while v.any():
            par ^= v & np.uint64(1)
            v >>= np.uint64(1)
        return np.where(par == 1, self.soft_decision_high, self.soft_decision_low).astype(_soft_dtype(self.soft_bytes))


class ViterbiDecoder_CUDA:
    """ViterbiDecoder_Core + the stateless Decoder::update rolled into one object backed by GPU state."""

    def __init__(self, branch_table: ViterbiBranchTable, config: ViterbiDecoder_Config, device=0, tie_break=VITB_TIE_SCALAR):
        self._L = _lib.load()
        self.branch_table, self.config = branch_table, config
        self.K, self.R = branch_table.K, branch_table.R
        self.NUMSTATES = 1 << (self.K - 1)
        self.soft_bytes = branch_table.soft_bytes
        p = vitb_params()
        p.K, p.R = self.K, self.R
        for i, g in enumerate(branch_table.G):
            p.G[i] = g
        p.soft_bytes = self.soft_bytes
        p.soft_decision_high, p.soft_decision_low = branch_table.soft_decision_high, branch_table.soft_decision_low
        p.soft_decision_max_error = config.soft_decision_max_error
        p.initial_start_error = config.initial_start_error
        p.initial_non_start_error = config.initial_non_start_error
        p.renormalisation_threshold = config.renormalisation_threshold
        p.tie_break, p.device = tie_break, device
        self._params = p
        self._h = C.c_void_p()
        _check(self._L.vitb_create(C.byref(p), C.byref(self._h)), "vitb_create")

    # -- lifetime ------------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.vitb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def is_valid(branch_table, config, tie_break=VITB_TIE_SCALAR):
        """static constexpr bool Decoder::is_valid (scalar.h:25)"""
        p = vitb_params()
        p.K, p.R = branch_table.K, branch_table.R
        for i, g in enumerate(branch_table.G):
            p.G[i] = g
        p.soft_bytes = branch_table.soft_bytes
        p.soft_decision_high, p.soft_decision_low = branch_table.soft_decision_high, branch_table.soft_decision_low
        p.soft_decision_max_error = config.soft_decision_max_error
        p.initial_start_error, p.initial_non_start_error = config.initial_start_error, config.initial_non_start_error
        p.renormalisation_threshold = config.renormalisation_threshold
        p.tie_break = tie_break
        return bool(_lib.load().vitb_is_supported(C.byref(p)))

    # -- reference call protocol ------------------------------------------------------------------------------------------
    def set_traceback_length(self, traceback_length):
        _check(self._L.vitb_set_traceback_length(self._h, traceback_length), "set_traceback_length")

    def get_traceback_length(self):
        n = C.c_size_t()
        _check(self._L.vitb_get_traceback_length(self._h, C.byref(n)))
        return n.value

    def reset(self, starting_state=0):
        _check(self._L.vitb_reset(self._h, starting_state), "reset")

    def update(self, symbols):
        """returns the sum of renormalisation minima of this call (what Decoder::update returns)"""
        s = np.ascontiguousarray(symbols, dtype=_soft_dtype(self.soft_bytes))
        acc = C.c_uint64()
        _check(self._L.vitb_update(self._h, s.ctypes.data, s.size, C.byref(acc)), "update")
        return acc.value

    def get_error(self, end_state=0):
        e = C.c_uint32()
        _check(self._L.vitb_get_error(self._h, end_state, C.byref(e)), "get_error")
        return e.value

    def chainback(self, total_bits, end_state=0):
        out = np.zeros((total_bits + 7) // 8, dtype=np.uint8)
        _check(self._L.vitb_chainback(self._h, out.ctypes.data, total_bits, end_state), "chainback")
        return out

    # -- public fields of the reference Core ---------------------------------------------------------------------------------
    @property
    def m_current_decoded_bit(self):
        n = C.c_size_t()
        _check(self._L.vitb_get_current_decoded_bit(self._h, C.byref(n)))
        return n.value

    @property
    def m_metrics(self):
        out = np.zeros(self.NUMSTATES, dtype=np.uint32)
        _check(self._L.vitb_get_metrics(self._h, out.ctypes.data))
        return out

    def m_decisions(self, first_row=0, n_rows=None):
        if n_rows is None:
            n_rows = self.m_current_decoded_bit - first_row
        words = max(self.NUMSTATES // 64, 1)
        out = np.zeros((n_rows, words), dtype=np.uint64)
        _check(self._L.vitb_get_decisions(self._h, first_row, n_rows, out.ctypes.data))
        return out

    # -- batched entry points ----------------------------------------------------------------------------------------------
    def set_puncture_schedule(self, keep, unpunctured_value=0):
        if keep is None or len(keep) == 0:
            _check(self._L.vitb_set_puncture_schedule(self._h, None, 0, 0))
            return
        k = np.ascontiguousarray(keep, dtype=np.uint8)
        _check(self._L.vitb_set_puncture_schedule(self._h, k.ctypes.data, k.size, unpunctured_value), "set_puncture_schedule")

    def _opts(self, row_stride, starting_state, end_state):
        o = vitb_batch_opts()
        o.row_stride, o.starting_state, o.end_state = row_stride, starting_state, end_state
        return o

    def decode_batch(self, symbols, total_bits, starting_state=0, end_state=0, want=("bytes", "acc", "final")):
        """symbols: host array [n_frames, row] of soft_t.  Returns (bytes [F, ceil(L/8)], acc_error [F], final_error [F])."""
        s = np.ascontiguousarray(symbols, dtype=_soft_dtype(self.soft_bytes))
        assert s.ndim == 2
        F = s.shape[0]
        out = np.zeros((F, (total_bits + 7) // 8), dtype=np.uint8) if "bytes" in want else None
        acc = np.zeros(F, dtype=np.uint64) if "acc" in want else None
        fin = np.zeros(F, dtype=np.uint32) if "final" in want else None
        o = self._opts(s.shape[1], starting_state, end_state)
        _check(self._L.vitb_decode_batch(self._h, s.ctypes.data, F, total_bits, C.byref(o),
                                         out.ctypes.data if out is not None else None,
                                         acc.ctypes.data if acc is not None else None,
                                         fin.ctypes.data if fin is not None else None), "decode_batch")
        return out, acc, fin

    def decode_batch_dev(self, d_symbols, n_frames, total_bits, d_out, d_acc, d_final, stream=0, row_stride=0, starting_state=0, end_state=0):
        """raw device pointers (ints), asynchronous on `stream` (a cudaStream_t as int)"""
        o = self._opts(row_stride, starting_state, end_state)
        _check(self._L.vitb_decode_batch_dev(self._h, d_symbols, n_frames, total_bits, C.byref(o), d_out, d_acc, d_final, stream),
               "decode_batch_dev")

    def decode_batch_async(self, h_symbols, n_frames, total_bits, h_out, h_acc, h_final, stream=0, row_stride=0, starting_state=0, end_state=0):
        """raw (pinned) HOST pointers (ints); H2D, kernels and D2H are enqueued on `stream`"""
        o = self._opts(row_stride, starting_state, end_state)
        _check(self._L.vitb_decode_batch_async(self._h, h_symbols, n_frames, total_bits, C.byref(o), h_out, h_acc, h_final, stream),
               "decode_batch_async")

    def set_traceback_window(self, window_bits=0):
        """vitb_set_traceback_window: K = 15 only; 0 = keep every decision row (exact mode)"""
        _check(self._L.vitb_set_traceback_window(self._h, window_bits), "set_traceback_window")

    @property
    def window_mismatches(self):
        n = C.c_uint64(0)
        _check(self._L.vitb_get_window_mismatches(self._h, C.byref(n)))
        return int(n.value)

    def set_pipelining(self, enabled=True):
        """vitb_set_pipelining: results of a decode_batch_dev call are complete (in stream order) after the NEXT call or batch_flush"""
        _check(self._L.vitb_set_pipelining(self._h, 1 if enabled else 0))

    def batch_flush(self, stream=0):
        _check(self._L.vitb_batch_flush(self._h, C.c_void_p(stream)))

    def set_profiling(self, enabled=True):
        _check(self._L.vitb_set_profiling(self._h, 1 if enabled else 0))

    def stage_ms(self):
        """device ms of the last batch call: {ingest, acs, traceback, gather}; synchronise the stream first"""
        ms = (C.c_float * 4)()
        _check(self._L.vitb_get_stage_ms(self._h, C.byref(ms)))
        return {"ingest": ms[0], "acs": ms[1], "traceback": ms[2], "gather": ms[3]}

    # -- device-side front end (include/viterbi_b200.h: vitb_synth_frames / vitb_quantise / vitb_ber_trial) ------------------
    def synth_frames(self, n_frames, total_bits, EbNo_dB=None, seed=1, row=None):
        """-> (tx_bytes [F, L/8], symbols [F, row]) generated on the GPU; EbNo_dB None = noise free"""
        n_sym = (total_bits + self.K - 1) * self.R
        row = row or n_sym
        tx = np.zeros((n_frames, total_bits // 8), dtype=np.uint8)
        sym = np.zeros((n_frames, row), dtype=_soft_dtype(self.soft_bytes))
        e = float("nan") if EbNo_dB is None else float(EbNo_dB)
        _check(self._L.vitb_synth_frames(self._h, n_frames, total_bits, e, seed, tx.ctypes.data, sym.ctypes.data, row), "synth_frames")
        return tx, sym

    def synth_frames_dev(self, n_frames, total_bits, d_tx, d_symbols, EbNo_dB=None, seed=1, row_stride=0, stream=0):
        """device pointers (ints): [F][L/8] transmitted bytes and [F][row_stride] soft symbols are generated in place on `stream`"""
        e = float("nan") if EbNo_dB is None else float(EbNo_dB)
        _check(self._L.vitb_synth_frames_dev(self._h, n_frames, total_bits, e, seed, d_tx, d_symbols, row_stride, stream), "synth_frames_dev")

    def quantise(self, x, scale, mean):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros(x.shape, dtype=_soft_dtype(self.soft_bytes))
        _check(self._L.vitb_quantise(self._h, x.ctypes.data, x.size, scale, mean, out.ctypes.data), "quantise")
        return out

    def ber_trial(self, n_frames, total_bits, EbNo_dB, seed=1):
        """bit errors of n_frames frames generated, decoded and counted on the device"""
        c = C.c_uint64()
        e = float("nan") if EbNo_dB is None else float(EbNo_dB)
        _check(self._L.vitb_ber_trial(self._h, n_frames, total_bits, e, seed, C.byref(c)), "ber_trial")
        return c.value

    # -- introspection ---------------------------------------------------------------------------------------------------------
    def workspace_bytes(self, n_frames, total_bits):
        n = C.c_size_t()
        _check(self._L.vitb_workspace_bytes(self._h, n_frames, total_bits, C.byref(n)))
        return n.value

    def set_workspace_limit(self, nbytes):
        _check(self._L.vitb_set_workspace_limit(self._h, nbytes))

    def set_variant(self, lanes_per_pair=0):
        """pin the lanes-per-frame-pair kernel variant for batch calls (0 = choose by batch size)"""
        _check(self._L.vitb_set_variant(self._h, lanes_per_pair), "set_variant")

    def set_history_kernel(self, enabled=True):
        """batch calls of one-lane-per-pair variants: survivor-history kernel (default) or the decision-row kernels"""
        _check(self._L.vitb_set_history_kernel(self._h, 1 if enabled else 0), "set_history_kernel")

    def set_traceback_segments(self, seg_records=0, overlap_records=-1):
        """segment length / warm-up of the history-record traceback in records (0 / -1 = automatic)"""
        _check(self._L.vitb_set_traceback_segments(self._h, seg_records, overlap_records), "set_traceback_segments")

    @property
    def variants(self):
        buf = (C.c_int * 16)()
        n = self._L.vitb_get_variants(self._h, buf, 16)
        return [buf[i] for i in range(n)]

    @property
    def kernel_launch_count(self):
        n = C.c_uint64()
        _check(self._L.vitb_kernel_launch_count(self._h, C.byref(n)))
        return n.value

    @property
    def kernel_name(self):
        return self._L.vitb_kernel_name(self._h).decode()


def decode_batch_multi(decoders, symbols, total_bits, starting_state=0, end_state=0):
    """frames split into contiguous ranges, one per decoder (each on its own GPU), no collective (include/viterbi_b200.h)"""
    L = _lib.load()
    d0 = decoders[0]
    s = np.ascontiguousarray(symbols, dtype=_soft_dtype(d0.soft_bytes))
    F = s.shape[0]
    out = np.zeros((F, (total_bits + 7) // 8), dtype=np.uint8)
    acc = np.zeros(F, dtype=np.uint64)
    fin = np.zeros(F, dtype=np.uint32)
    hs = (C.c_void_p * len(decoders))(*[d._h.value for d in decoders])
    o = vitb_batch_opts()
    o.row_stride, o.starting_state, o.end_state = s.shape[1], starting_state, end_state
    _check(L.vitb_decode_batch_multi(hs, len(decoders), s.ctypes.data, F, total_bits, C.byref(o), out.ctypes.data, acc.ctypes.data,
                                     fin.ctypes.data), "decode_batch_multi")
    return out, acc, fin


def decode_batch_multi_raw(decoders, h_symbols, n_frames, total_bits, h_out, h_acc, h_final, row_stride=0, starting_state=0, end_state=0):
    """vitb_decode_batch_multi on raw HOST pointers (ints; pinned memory lets one host thread drive every device asynchronously)"""
    L = _lib.load()
    hs = (C.c_void_p * len(decoders))(*[d._h.value for d in decoders])
    o = vitb_batch_opts()
    o.row_stride, o.starting_state, o.end_state = row_stride, starting_state, end_state
    _check(L.vitb_decode_batch_multi(hs, len(decoders), h_symbols, n_frames, total_bits, C.byref(o), h_out, h_acc, h_final), "decode_batch_multi")
