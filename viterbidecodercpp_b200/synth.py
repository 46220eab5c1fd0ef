"""Synthetic BPSK-over-AWGN soft symbols with the statistics of the reference's BER sweep (examples/run_snr_ber.cpp:311-359):
random bytes -> rate-1/R convolutional encoder with K-1 zero tail bits -> +-1.0 -> + N(0, sigma^2) with
EsNo_dB = EbNo_dB - 10 log10(R), sigma^2 = 10^(-(EsNo_dB+3)/10) -> scale by (high-low)/2 / sqrt(1+sigma^2), add (high+low)/2,
round, clamp to [low, high].  Vectorised numpy (input generation only; never part of a timed region)."""
import numpy as np


def conv_encode(K, R, G, data_bytes):
    """data_bytes [F, nbytes] uint8 -> code bits [F, (nbytes*8 + K-1), R] in {0,1}.
    Bits enter MSB-first; output i at time t = parity(G[i] & reg), reg = (reg << 1) | bit, so tap k of G[i] sees the bit k steps
    back (include/viterbi/convolutional_encoder_shift_register.h:45-61), followed by K-1 zero bits
    (examples/helpers/test_helpers.h:53-61)."""
    data_bytes = np.atleast_2d(np.asarray(data_bytes, dtype=np.uint8))
    F, nb = data_bytes.shape
    bits = np.unpackbits(data_bytes, axis=1)                       # MSB first
    S = nb * 8 + K - 1
    padded = np.zeros((F, S + K - 1), dtype=np.uint8)              # K-1 leading zeros (reg starts at 0) + data + K-1 tail zeros
    padded[:, K - 1:K - 1 + nb * 8] = bits
    out = np.zeros((F, S, R), dtype=np.uint8)
    for i in range(R):
        acc = np.zeros((F, S), dtype=np.uint8)
        for k in range(K):
            if (G[i] >> k) & 1:
                acc ^= padded[:, K - 1 - k:K - 1 - k + S]
        out[:, :, i] = acc
    return out


def awgn_soft_symbols(code_bits, R, high, low, EbNo_dB, rng, dtype):
    """code bits {0,1} (any shape) -> quantised soft symbols, following run_snr_ber.cpp:320-359"""
    x = np.where(code_bits > 0, np.float32(1.0), np.float32(-1.0))
    if EbNo_dB is not None:
        EsNo_dB = EbNo_dB - 10.0 * np.log10(float(R))
        noise_variance = 10.0 ** (-(EsNo_dB + 3.0) / 10.0)
        norm = 1.0 / np.sqrt(1.0 + noise_variance)
        x = x + rng.normal(0.0, np.sqrt(noise_variance), size=x.shape).astype(np.float32)
    else:
        norm = 1.0
    mean = (high + low) / 2.0
    mag = (high - low) / 2.0
    y = np.round(x * np.float32(mag * norm) + np.float32(mean))
    return np.clip(y, low, high).astype(dtype)


def make_frames(K, R, G, n_frames, total_bits, high, low, soft_bytes, EbNo_dB, seed):
    """-> (tx_bytes [F, L/8], symbols [F, (L+K-1)*R] soft_t)"""
    assert total_bits % 8 == 0
    rng = np.random.default_rng(seed)
    tx = rng.integers(0, 256, size=(n_frames, total_bits // 8), dtype=np.uint8)
    bits = conv_encode(K, R, G, tx)
    sym = awgn_soft_symbols(bits, R, high, low, EbNo_dB, rng, np.int8 if soft_bytes == 1 else np.int16)
    return tx, sym.reshape(n_frames, -1)


def puncture(symbols, keep):
    """drop the symbols whose keep[] entry is 0 (transmit side of examples/helpers/puncture_code_helpers.h:57-98)"""
    keep = np.asarray(keep, dtype=bool)
    return np.ascontiguousarray(symbols[:, keep])
