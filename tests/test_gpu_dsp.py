"""GPU parity of the DSP kernels (through the C ABI) against the oracle / torch fp64."""
import numpy as np
import pytest
import torch

from oracle import dsp, synth

pytestmark = pytest.mark.gpu

GEOMS = list(dsp.GEOMETRIES.values())


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _ops():
    import se_b200
    return se_b200.ops


def ref_spec(wav64, geom):
    n_fft, win, hop = geom
    return np.stack([dsp.stft(w, n_fft, win, hop, out_dtype=np.complex128) for w in wav64])   # [B,F,T]


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("n", [16000, 16000 + 74, 2048])
def test_rms_and_stft_complex(geom, n):
    dev, ops = _dev(), _ops()
    n_fft, win, hop = geom
    b = 3
    wav = synth.noisy_batch(b, n)
    w = torch.from_numpy(wav).to(dev)
    c, ic = ops.rms_scale(w)
    cref = np.sqrt(n / np.sum(wav.astype(np.float64) ** 2, axis=1))
    assert np.allclose(c.cpu().numpy(), cref, rtol=2e-7)
    assert np.allclose(ic.cpu().numpy(), 1.0 / cref, rtol=2e-7)
    t, f = 1 + n // hop, n_fft // 2 + 1
    spec = torch.empty(b, t, f, 2, device=dev)
    ops.stft(w, c, n_fft, win, hop, re=spec[..., 0], im=spec[..., 1])
    torch.cuda.synchronize()
    ref = ref_spec(wav.astype(np.float64) * cref[:, None], geom).transpose(0, 2, 1)
    got = torch.view_as_complex(spec).cpu().numpy()
    err = np.abs(got - ref).max()
    scale = np.abs(ref).max()
    print(f"stft {geom} n={n}: max err {err:.3e} (max |X| {scale:.1f})")
    assert err < 3e-5 * max(1.0, scale / 50)


@pytest.mark.parametrize("geom", [GEOMS[1], GEOMS[2]])
def test_stft_freq_major_compressed_planes(geom):
    """[B,2,F,T] compressed RI (DCCRN/dccrn_decode.py:42-46) and [B,1,F,T] magnitude
    (fullsubnet_sa_decode.py:57-61)."""
    dev, ops = _dev(), _ops()
    n_fft, win, hop = geom
    b, n = 2, 12800
    wav = synth.noisy_batch(b, n, first_index=5)
    w = torch.from_numpy(wav).to(dev)
    c, _ = ops.rms_scale(w)
    t, f = 1 + n // hop, n_fft // 2 + 1
    ri = torch.empty(b, 2, f, t, device=dev)
    mag = torch.empty(b, f, t, device=dev)
    ops.stft(w, c, n_fft, win, hop, mag=None, re=ri[:, 0], im=ri[:, 1], layout="bft", p_ri=0.5)
    ops.stft(w, c, n_fft, win, hop, mag=mag, layout="bft", p_mag=0.5)
    cref = c.cpu().numpy().astype(np.float64)
    ref = ref_spec(wav.astype(np.float64) * cref[:, None], geom)
    m = np.abs(ref) ** 0.5
    ph = np.angle(ref)
    assert np.abs(mag.cpu().numpy() - m).max() < 2e-5
    assert np.abs(ri[:, 0].cpu().numpy() - m * np.cos(ph)).max() < 2e-5
    assert np.abs(ri[:, 1].cpu().numpy() - m * np.sin(ph)).max() < 2e-5


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("layout", ["btf", "bft"])
def test_istft_spec_matches_oracle(geom, layout):
    dev, ops = _dev(), _ops()
    from se_b200._lib import ISTFT_SPEC
    n_fft, win, hop = geom
    b, n = 2, 16000
    rng = np.random.default_rng(0)
    t, f = 1 + n // hop, n_fft // 2 + 1
    spec = (rng.standard_normal((b, f, t)) + 1j * rng.standard_normal((b, f, t))).astype(np.complex64)
    ref = np.stack([dsp.istft(s, n_fft, win, hop, length=n) for s in spec])
    re = torch.from_numpy(np.ascontiguousarray(spec.real)).to(dev)
    im = torch.from_numpy(np.ascontiguousarray(spec.imag)).to(dev)
    if layout == "btf":
        re, im = re.transpose(1, 2).contiguous(), im.transpose(1, 2).contiguous()
    out = torch.empty(b, n, device=dev)
    ops.istft(ISTFT_SPEC, re, im, None, None, n_fft, win, hop, out, n, layout_a=layout)
    err = np.abs(out.cpu().numpy() - ref).max()
    print(f"istft {geom} {layout}: max err {err:.3e}")
    assert err < 2e-5
    # length rule: longer than the overlap-add extent -> zero padded (librosa fix_length)
    out2 = torch.full((b, n + 500), 7.0, device=dev)
    ops.istft(ISTFT_SPEC, re, im, None, None, n_fft, win, hop, out2, n + 500, layout_a=layout)
    ref2 = np.stack([dsp.istft(s, n_fft, win, hop, length=n + 500) for s in spec])
    # the last n_fft/2 samples divide by a vanishing window envelope (values ~1e4): relative gate
    d2 = np.abs(out2.cpu().numpy() - ref2)
    assert (d2 <= 2e-5 * np.maximum(1.0, np.abs(ref2))).all()
    assert np.all(out2.cpu().numpy()[:, n + n_fft // 2:] == 0)


def test_istft_prologue_modes():
    dev, ops = _dev(), _ops()
    from se_b200._lib import ISTFT_CMASK, ISTFT_MAG_PHASE, ISTFT_RI_DECOMP
    n_fft, win, hop = 512, 512, 256
    b, n = 2, 8192
    t, f = 1 + n // hop, n_fft // 2 + 1
    rng = np.random.default_rng(1)
    X = (rng.standard_normal((b, f, t)) + 1j * rng.standard_normal((b, f, t))).astype(np.complex64)
    X[0, 3, 4] = 0.0                                            # angle(0) = 0 convention
    A = (rng.standard_normal((b, f, t)) + 1j * rng.standard_normal((b, f, t))).astype(np.complex64)
    est = np.abs(rng.standard_normal((b, f, t))).astype(np.float32)
    inv_scale = torch.tensor([0.5, 2.0], device=dev)

    def dev_planes(z):
        return (torch.from_numpy(np.ascontiguousarray(z.real)).to(dev),
                torch.from_numpy(np.ascontiguousarray(z.imag)).to(dev))

    xr, xi = dev_planes(X)
    ar, ai = dev_planes(A)
    out = torch.empty(b, n, device=dev)
    # (i) magnitude x noisy phase, p = 0.5 -> est**2   (lstm_decode.py:44,51)
    ops.istft(ISTFT_MAG_PHASE, torch.from_numpy(est).to(dev), None, xr, xi, n_fft, win, hop, out, n,
              out_scale=inv_scale, inv_p=2.0, layout_a="bft", layout_b="bft")
    Y = (est.astype(np.float64) ** 2) * np.exp(1j * np.angle(X))
    ref = np.stack([dsp.istft(y, n_fft, win, hop, length=n) for y in Y]) * np.array([0.5, 2.0])[:, None]
    e1 = np.abs(out.cpu().numpy() - ref).max()
    # (ii) RI decompress p = 0.5  (dccrn_decode.py:49-54)
    ops.istft(ISTFT_RI_DECOMP, ar, ai, None, None, n_fft, win, hop, out, n, inv_p=2.0, layout_a="bft")
    Y = (np.abs(A).astype(np.float64) ** 2) * np.exp(1j * np.angle(A))
    ref = np.stack([dsp.istft(y, n_fft, win, hop, length=n) for y in Y])
    e2 = np.abs(out.cpu().numpy() - ref).max()
    # (iii) complex mask on the compressed spectrum, then decompress (fullsubnet_sa_decode.py:64-73)
    ops.istft(ISTFT_CMASK, ar, ai, xr, xi, n_fft, win, hop, out, n, inv_p=2.0, p_x=0.5, layout_a="bft",
              layout_b="bft")
    Xc = (np.abs(X).astype(np.float64) ** 0.5) * np.exp(1j * np.angle(X))
    Cc = A.astype(np.complex128) * Xc
    Y = (np.abs(Cc) ** 2) * np.exp(1j * np.angle(Cc))
    ref = np.stack([dsp.istft(y, n_fft, win, hop, length=n) for y in Y])
    e3 = np.abs(out.cpu().numpy() - ref).max()
    print(f"istft prologues: mag_phase {e1:.3e} ri_decomp {e2:.3e} cmask {e3:.3e}")
    assert e1 < 5e-5 and e2 < 5e-5 and e3 < 2e-4


@pytest.mark.parametrize("geom", GEOMS)
def test_full_size_roundtrip_and_linearity(geom):
    """BASELINE-size batch (64 x 4 s): STFT -> identity -> iSTFT reproduces the input, and the
    transform is linear in the waveform (size-independent properties)."""
    dev = _dev()
    import se_b200
    b, n = 64, 64000
    g = torch.Generator(device="cpu").manual_seed(0)
    wav = (torch.rand(b, n, generator=g) - 0.5).to(dev)
    y = se_b200.decode.dsp_roundtrip(wav, geom)
    err = (y - wav).abs().max().item()
    print(f"roundtrip {geom}: max err {err:.3e}")
    assert err < 2e-5
    ops = se_b200.ops
    n_fft, win, hop = geom
    t, f = 1 + n // hop, n_fft // 2 + 1
    s1 = torch.empty(8, t, f, 2, device=dev)
    s2 = torch.empty_like(s1)
    s3 = torch.empty_like(s1)
    ops.stft(wav[:8], None, n_fft, win, hop, re=s1[..., 0], im=s1[..., 1])
    ops.stft(wav[8:16], None, n_fft, win, hop, re=s2[..., 0], im=s2[..., 1])
    ops.stft((wav[:8] + 2 * wav[8:16]).contiguous(), None, n_fft, win, hop, re=s3[..., 0], im=s3[..., 1])
    assert (s3 - (s1 + 2 * s2)).abs().max().item() < 5e-4


def test_dsp_rejects_bad_geometry():
    dev, ops = _dev(), _ops()
    import se_b200
    w = torch.zeros(1, 4000, device=dev)
    out = torch.empty(1, 26, 129, device=dev)
    with pytest.raises(se_b200._lib.SeB200Error):
        ops.stft(w, None, 256, 256, 160, mag=out)


@pytest.mark.parametrize("sr,n", [(48000, 48000 + 77), (44100, 30011), (8000, 9000), (32000, 4000)])
def test_resample_matches_oracle(sr, n):
    """se_resample (front step of the *_decode_vb.py scripts) vs the restated librosa/resampy algorithm.  fp32 table
    and accumulation against the float64 oracle: 1e-5 on unit-scale noise."""
    dev = _dev()
    import se_b200
    from oracle import resample as R
    rng = np.random.default_rng(sr + n)
    x = (rng.standard_normal((3, n)) * 0.3).astype(np.float32)
    y = se_b200.ops.resample(torch.from_numpy(x).to(dev), sr, 16000).cpu().numpy()
    for b in range(3):
        ref = R.librosa_resample(x[b].astype(np.float64), sr, 16000)
        assert y.shape[1] == len(ref)
        err = np.abs(y[b] - ref).max()
        assert err < 1e-5, (b, err)
    x16 = torch.from_numpy(x).to(dev)
    assert se_b200.ops.resample(x16, 16000, 16000) is x16


def test_resample_full_size_batch_properties():
    """64 x 4 s at 48 kHz: linearity and agreement with the oracle on two clips."""
    dev = _dev()
    import se_b200
    from oracle import resample as R, synth
    base = synth.noisy_batch(4, 64000)
    x = np.repeat(base, 3, axis=1).astype(np.float32)                  # crude 48 kHz material, 192 000 samples
    x = np.tile(x, (16, 1)) * np.linspace(0.5, 1.5, 64, dtype=np.float32)[:, None]
    xt = torch.from_numpy(x).to(dev)
    y = se_b200.ops.resample(xt, 48000, 16000)
    assert y.shape == (64, 64000)
    y2 = se_b200.ops.resample(2.0 * xt[:8] + xt[8:16], 48000, 16000)
    assert (y2 - (2.0 * y[:8] + y[8:16])).abs().max().item() < 1e-5
    for b in (0, 63):
        ref = R.librosa_resample(x[b].astype(np.float64), 48000, 16000)
        assert np.abs(y[b].cpu().numpy() - ref).max() < 1e-5



@pytest.mark.parametrize("geom", GEOMS)
def test_ragged_batch_equals_per_clip_dsp(geom):
    """Tail-padded batch with per-clip lengths (se_rms_scale_len / se_stft_len / se_istft_len): the frames, the scale and the
    reconstructed samples of every clip are BIT-identical to that clip processed alone at its own length; frames and
    samples past a clip's end are zero.  Lengths include non-multiples of the hop and the shortest legal clip."""
    dev, ops = _dev(), _ops()
    from se_b200._lib import ISTFT_SPEC
    n_fft, win, hop = geom
    lens = [n_fft, 3 * hop + n_fft + 1, 9000, 12345, 16000, 16000 - 1]
    nmax = max(lens)
    b = len(lens)
    wav = np.zeros((b, nmax), dtype=np.float32)
    for i, n in enumerate(lens):
        wav[i, :n] = synth.noisy_clip(50 + i, n)
    w = torch.from_numpy(wav).to(dev)
    lt = torch.tensor(lens, dtype=torch.int32, device=dev)
    tmax, f = 1 + nmax // hop, n_fft // 2 + 1
    c, ic = ops.rms_scale(w, lengths=lt)
    spec = torch.full((b, tmax, f, 2), 7.0, device=dev)
    mag = torch.full((b, tmax, f), 7.0, device=dev)
    ops.stft(w, c, n_fft, win, hop, mag=mag, re=spec[..., 0], im=spec[..., 1], lengths=lt)
    out = torch.full((b, nmax), 7.0, device=dev)
    ops.istft(ISTFT_SPEC, spec[..., 0], spec[..., 1], None, None, n_fft, win, hop, out, nmax, out_scale=ic, lengths=lt)
    torch.cuda.synchronize()
    for i, n in enumerate(lens):
        wi = w[i:i + 1, :n].contiguous()
        ti = 1 + n // hop
        c1, ic1 = ops.rms_scale(wi)
        s1 = torch.empty(1, ti, f, 2, device=dev)
        m1 = torch.empty(1, ti, f, device=dev)
        ops.stft(wi, c1, n_fft, win, hop, mag=m1, re=s1[..., 0], im=s1[..., 1])
        o1 = torch.empty(1, n, device=dev)
        ops.istft(ISTFT_SPEC, s1[..., 0], s1[..., 1], None, None, n_fft, win, hop, o1, n, out_scale=ic1)
        assert torch.equal(c[i:i + 1], c1) and torch.equal(ic[i:i + 1], ic1)
        assert torch.equal(spec[i, :ti], s1[0]) and torch.equal(mag[i, :ti], m1[0])
        assert float(spec[i, ti:].abs().sum()) == 0.0 and float(mag[i, ti:].abs().sum()) == 0.0
        assert torch.equal(out[i, :n], o1[0]) and float(out[i, n:].abs().sum()) == 0.0
        assert (o1[0] - wi[0]).abs().max().item() < 2e-6          # and it is the identity
