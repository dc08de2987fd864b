"""Frame generator (csrc/framegen.cuh): Philox known answers, encoders against the host mirrors
of the reference's encoders, noise statistics, shard independence."""
import numpy as np
import pytest

import polarcode_and_ldpc_b200 as P
from tests.emu import emu


def _philox_py(ctr, key):
    """Philox4x32-10 restated from the published algorithm (Salmon et al., SC'11, Random123)."""
    c = [int(x) for x in ctr]
    k = [int(x) for x in key]
    M0, M1, W0, W1, mask = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & mask, p1 & mask, ((p0 >> 32) ^ c[3] ^ k[1]) & mask, p0 & mask]
        k = [(k[0] + W0) & mask, (k[1] + W1) & mask]
    return np.array(c, dtype=np.uint32)


# Random123 known-answer vectors for philox4x32_10 (kat_vectors)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, out in KAT:
        assert tuple(int(x) for x in _philox_py(ctr, key)) == out
        assert tuple(int(x) for x in emu.philox(ctr, key)) == out
    rng = np.random.default_rng(0)
    for _ in range(50):
        ctr, key = rng.integers(0, 2 ** 32, 4), rng.integers(0, 2 ** 32, 2)
        assert np.array_equal(emu.philox(ctr, key), _philox_py(ctr, key))


def _check_frames(kind, N, K, table, enc, llr, msg, cw, snr_db, rtol):
    assert set(np.unique(msg)) <= {0, 1} and set(np.unique(cw)) <= {0, 1}
    assert np.array_equal(enc.encode_batch(msg.astype(np.int64)), cw.astype(np.int64)), "encoder mismatch"
    ch = P.AWGNChannel(snr_db)
    z = (llr.astype(np.float64) * ch.noise_std ** 2 / 2.0 - (1.0 - 2.0 * cw)) / ch.noise_std
    n = z.size
    assert abs(z.mean()) < 5.0 / np.sqrt(n)
    assert abs(z.var() - 1.0) < 5.0 * np.sqrt(2.0 / n) + rtol
    assert abs((z ** 4).mean() - 3.0) < 5.0 * np.sqrt(96.0 / n) + 10 * rtol
    assert abs(msg.mean() - 0.5) < 5.0 * 0.5 / np.sqrt(msg.size)


def test_emu_polar_and_ldpc_frames():
    for N, K in ((256, 100), (1024, 512), (16, 7), (64, 33), (2048, 1000)):
        fz = P.bhattacharyya_frozen_set(N, K, 2.0)
        F = max(8, 16384 // N)
        llr, msg, cw = emu.gen_frames("polar", N, K, fz, F, 1.5, seed=11)
        _check_frames("polar", N, K, fz, P.PolarEncoder(N, K, fz), llr, msg, cw, 1.5, 1e-3)
    for n in (96, 504):
        H = P.gallager_parity_check(n, 3, 6, 42)
        G, _ = P.generator_from_parity(H)
        llr, msg, cw = emu.gen_frames("ldpc", n, G.shape[0], G, 40, 0.5, seed=5, dtype="f64")
        _check_frames("ldpc", n, G.shape[0], G, P.LDPCEncoder(n, G.shape[0], H=H, G=G), llr, msg, cw, 0.5, 1e-3)
        assert not ((H @ cw.T.astype(np.int64)) % 2).any()


def test_emu_frames_do_not_depend_on_sharding():
    N, K = 128, 64
    fz = P.bhattacharyya_frozen_set(N, K, 2.0)
    full = emu.gen_frames("polar", N, K, fz, 37, 2.0, seed=99)
    lo = emu.gen_frames("polar", N, K, fz, 20, 2.0, seed=99, frame0=0)
    hi = emu.gen_frames("polar", N, K, fz, 17, 2.0, seed=99, frame0=20)
    for a, b, c in zip(full, lo, hi):
        assert np.array_equal(a, np.concatenate([b, c]))
    other = emu.gen_frames("polar", N, K, fz, 37, 2.0, seed=100)
    assert not np.array_equal(full[1], other[1])


def _check_other_channels(gen_fn):
    """Rayleigh: s LLR sigma^2 / 2 = h^2 + sigma h z with h^2 ~ Exp(1) (fading.py:38-47);
    BSC: flips with probability p, LLR = +-ln((1 - p) / p) (bsc.py:36-38)."""
    llr, _, cw = gen_fn(channel="rayleigh", param=3.0)
    sig = P.RayleighFadingChannel(3.0).noise_std
    v = (1.0 - 2.0 * cw) * llr.astype(np.float64) * sig ** 2 / 2.0
    n = v.size
    assert abs(v.mean() - 1.0) < 5.0 * np.sqrt((1.0 + sig ** 2) / n)              # E h^2 = 1
    assert abs(v.var() - (1.0 + sig ** 2)) < 0.05                                # Var h^2 + sigma^2 E h^2
    ref = P.RayleighFadingChannel(3.0)
    np.random.seed(0)
    host = (1.0 - 2.0 * cw) * ref.transmit_batch(cw) * sig ** 2 / 2.0
    assert abs(np.median(v) - np.median(host)) < 0.02 and abs((v < 0).mean() - (host < 0).mean()) < 0.01
    p = 0.07
    llr, _, cw = gen_fn(channel="bsc", param=p)
    mag = np.log((1 - p) / p)
    np.testing.assert_allclose(np.abs(llr), mag, rtol=1e-6)
    flips = (llr < 0).astype(np.int64) != cw
    assert abs(flips.mean() - p) < 5.0 * np.sqrt(p * (1 - p) / flips.size)


def test_emu_rayleigh_and_bsc_channels():
    fz = P.bhattacharyya_frozen_set(256, 128, 2.0)

    def gen_fn(channel, param):
        return emu.gen_frames("polar", 256, 128, fz, 96, param, seed=3, channel={"rayleigh": 1, "bsc": 2}[channel])
    _check_other_channels(gen_fn)


@pytest.mark.gpu
def test_gpu_rayleigh_and_bsc_channels():
    fz = P.bhattacharyya_frozen_set(1024, 512, 2.0)
    gen = P.FrameGenerator.polar(1024, 512, fz)

    def gen_fn(channel, param):
        llr, msg, cw = gen.generate(512, param, seed=3, channel=channel)
        return llr.cpu().numpy(), msg.cpu().numpy(), cw.cpu().numpy().astype(np.int64)
    _check_other_channels(gen_fn)
    # decoders take the LLRs of either channel: BSC(0.03) through SCL-8 at rate 1/2 decodes clean
    llr, msg, _ = gen.generate(2048, 0.03, seed=9, channel="bsc")
    c = P.count_errors(P.SCLDecoder(1024, 512, 8, fz).decode_batch(llr), msg)
    assert c[1].item() < 20


@pytest.mark.gpu
def test_gpu_framegen_matches_host_encoders_and_statistics():
    import torch
    for N, K in ((1024, 512), (256, 128), (4096, 3000), (32, 16)):
        fz = P.bhattacharyya_frozen_set(N, K, 2.0)
        gen = P.FrameGenerator.polar(N, K, fz)
        F = 4096 if N <= 1024 else 512
        llr, msg, cw = gen.generate(F, 2.0, seed=7)
        _check_frames("polar", N, K, fz, P.PolarEncoder(N, K, fz), llr.cpu().numpy(), msg.cpu().numpy(),
                      cw.cpu().numpy(), 2.0, 1e-4)
        # shard independence: frames [lo, hi) of a second call equal the slice of the first
        l2, m2, c2 = gen.generate(1000, 2.0, seed=7, frame0=96)
        assert torch.equal(l2[:400], llr[96:496]) and torch.equal(m2[:400], msg[96:496])
    for n in (504, 2016):
        H = P.gallager_parity_check(n, 3, 6, 42)
        G, _ = P.generator_from_parity(H)
        gen = P.FrameGenerator.ldpc(G)
        llr, msg, cw = gen.generate(2048, 1.0, seed=3, dtype="float64")
        _check_frames("ldpc", n, G.shape[0], G, P.LDPCEncoder(n, G.shape[0], H=H, G=G), llr.cpu().numpy(),
                      msg.cpu().numpy(), cw.cpu().numpy(), 1.0, 1e-4)
    from scipy import stats
    gen = P.FrameGenerator.polar(1024, 512, P.bhattacharyya_frozen_set(1024, 512, 2.0))
    llr, _, cw = gen.generate(2048, 0.0, seed=1)
    ch = P.AWGNChannel(0.0)
    z = ((llr.double() * ch.noise_std ** 2 / 2.0 - (1.0 - 2.0 * cw.double())) / ch.noise_std).cpu().numpy().ravel()
    assert stats.kstest(z[::7], "norm").pvalue > 1e-3
    assert abs((np.abs(z) > 3).mean() - 0.0026998) < 2e-4


@pytest.mark.gpu
def test_gpu_generated_frames_decode():
    """generator -> decoder -> error counters, all on the device: SCL-8 at 2 dB decodes clean
    and the emulator build produces the same frames as the CUDA build."""
    N, K = 1024, 512
    fz = P.bhattacharyya_frozen_set(N, K, 2.0)
    gen = P.FrameGenerator.polar(N, K, fz)
    llr, msg, cw = gen.generate(4096, 2.0, seed=5)
    bits = P.SCLDecoder(N, K, 8, fz).decode_batch(llr)          # CUDA tensor in -> CUDA tensor out
    counters = P.count_errors(bits, msg)
    assert counters[2].item() == 4096 and counters[1].item() < 41
    e_llr, e_msg, e_cw = emu.gen_frames("polar", N, K, fz, 8, 2.0, seed=5)
    assert np.array_equal(e_msg, msg[:8].cpu().numpy()) and np.array_equal(e_cw, cw[:8].cpu().numpy())
    np.testing.assert_allclose(e_llr, llr[:8].cpu().numpy(), rtol=2e-5, atol=2e-5)
