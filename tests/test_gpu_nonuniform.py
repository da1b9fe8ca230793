"""Non-uniformly partitioned Generic renderer (ssr_nu_*, csrc/nonuniform.cu; SURVEY 8f row 4 -- not in the
reference, so the checker is float64 direct convolution plus the reference's block-0 fade-in, and the uniform
GPU renderer, itself pinned to the reference).  Tolerance 1e-5 of full scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _truth(x, irs, gains, B):
    """x: (N, T) float input, irs[n]: (taps, M); the reference's GenericRenderer with constant weights:
    y[m] = sum_n g_n (x_n * h_nm), block 0 faded in from weight 0 (src/rendererbase.h:379 +
    apf/apf/combine_channels.h:297-335 with the raised cosine of apf/apf/math.h:254-259)."""
    N, T = x.shape
    M = irs[0].shape[1]
    y = np.zeros((M, T), np.float64)
    nfft = 1 << int(np.ceil(np.log2(T + max(h.shape[0] for h in irs))))
    for n in range(N):
        X = np.fft.rfft(x[n].astype(np.float64), nfft)
        for m in range(M):
            H = np.fft.rfft(irs[n][:, m].astype(np.float64), nfft)
            y[m] += gains[n] * np.fft.irfft(X * H, nfft)[:T]
    i = np.arange(B, dtype=np.float32)
    w = (np.cos(i * np.float32(2 * np.pi) / np.float32(2 * B)) * np.float32(0.5) + np.float32(0.5)).astype(np.float64)
    fade_in = np.concatenate([[0.0], w[:0:-1]])   # fade_in[i] = w[B - i], w[B] = 0
    y[:, :B] *= fade_in[None, :]
    return y


def _ir(rng, taps, M, scale):
    return (rng.uniform(-1, 1, (taps, M)) * np.exp(-4.0 * np.arange(taps) / taps)[:, None] * scale).astype(np.float32)


@pytest.mark.parametrize("B,levels,taps,M,N", [
    (64, [(1, 8), (4, 5)], 1700, 4, 3),             # two levels, 4 sub-renderers of one output each
    (64, [(1, 8), (2, 4), (8, 3)], 2500, 6, 2),     # three levels; outputs do not divide by 8
    (64, [(1, 4), (2, 6)], 900, 1, 2),              # fewer outputs than sub-renderer slots
    (128, [(1, 8), (4, 3)], 1300, 3, 2),            # IR ends inside the first tail partition
    (128, [(1, 8), (4, 3)], 1000, 3, 2),            # IR ends before the tail level (silent tail)
    (1024, [(1, 8), (4, 2)], 14000, 4, 2),          # the bulk-copy MAC at both levels (B = 1024 / 4096)
    (1024, [(1, 4), (2, 2), (4, 2), (8, 2)], 30000, 8, 2),   # four levels up to 8192-sample partitions
])
def test_nonuniform_equals_direct_convolution(capi, B, levels, taps, M, N):
    rng = np.random.default_rng(B + taps)
    ren = capi.NuRenderer(B, M, levels)
    assert ren.taps_covered >= taps
    scale = 0.5 / np.sqrt(N * taps / 8.0)
    irs = [_ir(rng, taps, M, scale) for _ in range(N)]
    ids = [ren.add_source(h) for h in irs]
    gains = [1.0] * N
    gains[0] = 0.7
    ren.set_gain(ids[0], 0.7)
    kmax = max(k for k, _ in levels)
    nblocks = 3 * kmax + (taps + B - 1) // B + 5
    x = rng.uniform(-1, 1, (N, nblocks * B)).astype(np.float32)
    y = np.concatenate([ren.process(x[:, t * B:(t + 1) * B]) for t in range(nblocks)], axis=1)
    ref = _truth(x, irs, gains, B)
    err = np.abs(y - ref).max()
    assert err <= TOL, f"non-uniform renderer deviates from float64 convolution by {err:.3e}"
    ren.close()


def test_nonuniform_matches_the_uniform_renderer_on_synthetic_sources(capi, engine_factory):
    """Same seeds -> same impulse responses (the tail levels generate taps [tap_first, ...) of them):
    the non-uniform renderer and the uniform one agree, outputs sharded or not."""
    import ssr_b200.workloads as wl
    B, M, N, taps = 256, 8, 3, 5000
    env = wl.envelope(taps)
    scale = 0.05
    eng = engine_factory(B)
    uni = capi.Renderer(eng, capi.GENERIC, outputs=M)
    nu = capi.NuRenderer(B, M, [(1, 8), (4, 3)])
    nu_shard = capi.NuRenderer(B, M, [(1, 8), (4, 3)], output_first=2, output_count=5)
    for n in range(N):
        for r in (uni, nu, nu_shard):
            r.add_source_synthetic(taps, M, 77, channel_id_offset=n * M, envelope=env, scale=scale)
    rng = np.random.default_rng(5)
    err = err_s = 0.0
    for t in range(40):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        a = uni.process(x)
        err = max(err, float(np.abs(nu.process(x) - a).max()))
        err_s = max(err_s, float(np.abs(nu_shard.process(x) - a[2:7]).max()))
    assert err <= TOL and err_s <= TOL, (err, err_s)
    assert nu_shard.local_outputs == 5
    nu.close(); nu_shard.close()


def test_nonuniform_mute_reaches_every_level(capi):
    """A mute ends a source everywhere: head at once (the reference's one-block fade-out), tail levels at
    their own block rate -- after the longest level's period plus the filter length the output is silent."""
    B, M, taps = 64, 2, 1500
    levels = [(1, 8), (4, 4)]
    rng = np.random.default_rng(9)
    ren = capi.NuRenderer(B, M, levels)
    sid = ren.add_source(_ir(rng, taps, M, 0.05))
    x = rng.uniform(-1, 1, (1, B)).astype(np.float32)
    for t in range(12):
        y = ren.process(x)
    assert np.abs(y).max() > 1e-3
    ren.set_mute(sid, True)
    for t in range(2 * 4 + (taps + B - 1) // B + 12):
        y = ren.process(x)
    assert np.abs(y).max() == 0.0
    ren.close()


def test_nonuniform_rejects_bad_level_layouts(capi):
    for levels in ([(2, 4)], [(1, 4), (4, 2)], [(1, 6), (4, 2)], [(1, 8), (4, 2), (4, 2)]):
        with pytest.raises(capi.SsrError):
            capi.NuRenderer(64, 2, levels)
    ren = capi.NuRenderer(64, 2, [(1, 8), (4, 2)])
    with pytest.raises(capi.SsrError):
        ren.add_source(np.zeros((8 * 64 + 2 * 256 + 1, 2), np.float32))   # longer than the levels cover
    with pytest.raises(capi.SsrError):
        ren.add_source(np.zeros((100, 3), np.float32))                    # wrong channel count
    ren.close()
