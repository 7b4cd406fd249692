"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run 2^30
samples): the recordings are device-resident and built by tiling one seeded block whose length is a multiple
of the hop, so that EVERY frame must reproduce frame 0 bit for bit (idempotence over time; 64-bit frame
indexing past 2^31 bytes), and frame 0 is checked against the oracle.  Parseval and a torch.fft (cuFFT, allowed
as a cross-check only) comparison on sampled frames cover the headline configuration, which has real structure
in every frame.  The general (any-alignment) kernels of the 2048..16384 sizes are exercised here as well.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from spectral_analyzer_b200 import synth, _capi
from util import check_db_parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def tiled_device_recording(block_bytes, repeats):
    blk = torch.from_numpy(np.ascontiguousarray(block_bytes)).to(DEV)
    return blk.repeat(repeats)


def run_device(engine, raw, datatype, nfft, hop, window, frames, out_kind="f32", **kw):
    obytes = {"f32": 4, "f64": 8, "rgba8": 4}[out_kind]
    out = torch.empty(frames * nfft * obytes, dtype=torch.uint8, device=DEV)
    p = engine.make_params(datatype, nfft, hop, window, n_frames=frames, out=out_kind, **kw)
    engine.spectrogram_device(raw.data_ptr(), raw.numel(), p, out.data_ptr(), out.numel(),
                              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out


def rows_all_equal_first(out_u8, row_bytes):
    rows = out_u8.view(-1, row_bytes)
    # compared in 64 MiB slabs to bound the temporary
    step = max(1, (64 << 20) // row_bytes)
    for i in range(0, rows.shape[0], step):
        if not torch.equal(rows[i:i + step], rows[0].expand_as(rows[i:i + step])):
            return False
    return True


def test_c1_headline_full_size_parseval_and_cufft_crosscheck(engine):
    """cf32, 2^28 samples (2 GiB), 1024-pt Hann, hop 512: the bench workload."""
    n, nfft, hop = 1 << 28, 1024, 512
    from bench import make_device_recording
    d_iq = make_device_recording(torch, 0, n, torch.device(DEV))
    frames = (n - nfft) // hop + 1
    out = run_device(engine, d_iq.view(torch.uint8), "cf32_le", nfft, hop, "hann", frames).view(torch.float32).view(frames, nfft)
    assert torch.isfinite(out).all()
    x = torch.view_as_complex(d_iq.view(-1, 2))
    w = torch.from_numpy(0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nfft) / nfft)).to(DEV)
    # (i) Parseval over ALL frames, 8192 frames at a time: sum_k |X_k|^2 = N * sum_n |w x|^2
    worst = 0.0
    for f0 in range(0, frames, 8192):
        f1 = min(frames, f0 + 8192)
        seg = x[f0 * hop: (f1 - 1) * hop + nfft].unfold(0, nfft, hop).to(torch.complex128) * w
        e = (seg.abs() ** 2).sum(dim=1) * nfft
        p = torch.pow(10.0, out[f0:f1].double() / 10.0).sum(dim=1)
        worst = max(worst, float((p / e - 1).abs().max()))
    assert worst < 2e-5, worst
    # (ii) sampled frames against cuFFT in FP64 (cross-check only), first / last / random, fft-shifted rows
    idx = torch.cat([torch.tensor([0, 1, frames - 2, frames - 1]), torch.randint(0, frames, (60,), generator=torch.Generator().manual_seed(5))])
    seg = torch.stack([x[int(i) * hop: int(i) * hop + nfft] for i in idx]).to(torch.complex128) * w
    ref = 20 * torch.log10(torch.fft.fftshift(torch.fft.fft(seg, dim=1), dim=1).abs() + 1e-10)
    check_db_parity(out[idx.to(DEV)].cpu().numpy(), ref.cpu().numpy())
    # (iii) time-sharding exactness at large offsets: rows of the second half recomputed with start_sample
    half = frames // 2
    p2 = engine.make_params("cf32_le", nfft, hop, "hann", n_frames=frames - half, start_sample=half * hop)
    out2 = torch.empty((frames - half, nfft), dtype=torch.float32, device=DEV)
    engine.spectrogram_device(d_iq.data_ptr(), d_iq.numel() * 4, p2, out2.data_ptr(), out2.numel() * 4,
                              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(out2, out[half:])


def test_c2_full_size_ci16_4096_blackman_harris(engine):
    """cs16, 2^30 samples (4 GiB in, 4 GiB out), 4096-pt Blackman-Harris, reference framing."""
    nfft = 4096
    blk = synth.recording(nfft, "ci16_le", seed=2)
    raw = tiled_device_recording(blk, (1 << 30) // nfft)
    frames = (1 << 30) // nfft
    out = run_device(engine, raw, "ci16_le", nfft, nfft, "blackman_harris", frames)
    assert rows_all_equal_first(out, nfft * 4)
    row0 = out[: nfft * 4].view(torch.float32).cpu().numpy()[None]
    check_db_parity(row0, co.spectrogram(blk, "ci16_le", 0, nfft, nfft, "blackman_harris", 1))
    # 50 % overlap over the same bytes: frames alternate between two rows (block period = 2 hops)
    frames2 = min(((1 << 30) - nfft) // 2048 + 1, 1 << 18)
    out2 = run_device(engine, raw, "ci16_le", nfft, 2048, "blackman_harris", frames2).view(frames2, nfft * 4)
    assert torch.equal(out2[0::2], out2[0].expand_as(out2[0::2])) and torch.equal(out2[1::2], out2[1].expand_as(out2[1::2]))
    ref2 = co.spectrogram(np.tile(blk, 2), "ci16_le", 0, nfft, 2048, "blackman_harris", 2)
    check_db_parity(out2[:2].view(torch.float32).cpu().numpy().reshape(2, nfft), ref2)


def test_c4_full_size_cu8_2048_rgba(engine):
    """cu8, 2^31 samples (one GPU's 4 GiB slice of the 16 GiB recording), 2048-pt, RGBA heatmap."""
    nfft, fs = 2048, 2.4e6
    blk = synth.recording(nfft, "cu8", seed=4)
    frames = (1 << 31) // nfft
    raw = tiled_device_recording(blk, frames)
    out = run_device(engine, raw, "cu8", nfft, nfft, "rect", frames, out_kind="rgba8", colormap="Heatmap", sample_rate=fs)
    assert rows_all_equal_first(out, nfft * 4)
    db = co.spectrogram(blk, "cu8", 0, nfft, nfft, "rect", 1)
    ref = co.render_rgba(db, fs, -160.0, -30.0, "Heatmap").astype(np.int32)
    got = out[: nfft * 4].cpu().numpy().reshape(1, nfft, 4).astype(np.int32)
    conv = 10 * np.log10(fs / nfft) + 20 * np.log10(nfft)
    nlev = np.clip((db - conv + 160.0) / 130.0, 0, 1)
    safe = (np.abs(nlev - 0.2) > 1e-4) & (np.abs(nlev - 0.5) > 1e-4)
    assert np.abs(got - ref)[safe].max() <= 1


def test_c5_full_size_cf64_65536(engine):
    """cf64, 2^26 samples (1 GiB), 65536-pt Hann, FP64 path, f64 dB out."""
    nfft = 65536
    blk = synth.recording(nfft, "cf64_le", seed=5)
    frames = (1 << 26) // nfft
    raw = tiled_device_recording(blk, frames)
    out = run_device(engine, raw, "cf64_le", nfft, nfft, "hann", frames, out_kind="f64")
    assert rows_all_equal_first(out, nfft * 8)
    row0 = out[: nfft * 8].view(torch.float64).cpu().numpy()[None]
    ref = co.spectrogram(blk, "cf64_le", 0, nfft, nfft, "hann", 1)
    check_db_parity(row0, ref, strong_tol=1e-9, floor_tol=1e-7)


def test_c3_full_size_annotation_batch(engine):
    """500 annotations x 2^20 samples, down 16, Welch 8192 / 75 % overlap on a 2^29-sample cf32 recording whose
    period (2^16) divides every annotation start and makes the NCO phase repeat: all 500 results must be
    identical bit for bit, and the first is checked against the oracle."""
    period, n, count, down, n_ann = 1 << 16, 1 << 29, 1 << 20, 16, 500
    blk = synth.recording(period, "cf32_le", seed=3)
    raw = tiled_device_recording(blk, n // period)
    f_off = 1024.0 / period                                  # integer number of cycles per period
    anns = (_capi.Annotation * n_ann)()
    offs = (C.c_uint64 * n_ann)()
    m = count // down
    rng = np.random.default_rng(3)
    for i in range(n_ann):
        start = int(rng.integers(0, (n - count) // period)) * period
        anns[i] = _capi.Annotation(start, count, f_off, down, 0)
        offs[i] = i * 2 * m
    d_iq = torch.empty(n_ann * 2 * m, dtype=torch.float64, device=DEV)
    d_psd = torch.empty(n_ann * 8192, dtype=torch.float64, device=DEV)
    _capi.check(_capi.lib().sa_downconvert_psd_batch_device(
        engine.handle, raw.data_ptr(), raw.numel(), 0, 0, 1.0e6, anns, n_ann, 8192, 2048, 1, d_iq.data_ptr(), offs,
        d_psd.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    iq = d_iq.view(n_ann, 2 * m)
    psd = d_psd.view(n_ann, 8192)
    assert torch.equal(iq, iq[0].expand_as(iq)) and torch.equal(psd, psd[0].expand_as(psd))
    host = np.tile(blk, count // period + 1)
    ref = co.downconvert(host, "cf32_le", 0, count, f_off, down, False)
    got = iq[0].view(2, m).cpu().numpy()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5
    ref_psd = co.psd_welch(ref, 1.0e6 / down, 8192)[1]
    # line spectrum (periodic input): compare the bins within 40 dB of the peak, the 'signal bins' of the tolerance
    top = ref_psd > ref_psd.max() - 40
    assert np.abs(psd[0].cpu().numpy() - ref_psd)[top].max() < 2e-3


@pytest.mark.parametrize("dt", ["cf32_le", "ci16_be", "cu8", "ci8"])
@pytest.mark.parametrize("nfft", [2048, 4096, 8192, 16384])
def test_general_and_aligned_kernels_agree_with_oracle(engine, dt, nfft):
    """Frames whose byte stride is a multiple of 16 take the small-radix-first kernels; an odd hop (stride 8, 4 or
    2 bytes times an odd number) takes the any-alignment kernels.  Both against the oracle."""
    frames = 7
    for hop in (nfft // 2, nfft // 2 + 1):
        raw = synth.recording((frames - 1) * hop + nfft + 16, dt, seed=nfft)
        ref = co.spectrogram(raw, dt, 5, nfft, hop, "hann", frames)
        got = engine.spectrogram(raw, dt, nfft, frames, hop=hop, window="hann", start_sample=5)
        check_db_parity(got, ref)
