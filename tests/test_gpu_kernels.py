"""Parity of every CUDA kernel (through the C ABI) against a plain PyTorch fp32 evaluation of the
same op on the same bf16-rounded operands.  Tolerances are stated per test: the kernels accumulate
in fp32, so against an fp32/fp64 evaluation of the SAME operands only summation order and the
bf16 rounding of outputs remain."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from cap4d_b200 import ops as _ops

    return _ops


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "M,N,K",
    [
        (128, 64, 64),        # one tile, one k-block
        (256, 128, 128),
        (1024, 320, 320),     # BN=160
        (384, 1280, 1280),    # many k-blocks, pipeline wrap-around
        (100, 96, 192),       # M tail (TMA zero fill + row mask)
        (32, 64, 128),        # M smaller than a tile
        (4096, 960, 320),     # QKV shape at C=320
        (2048, 2560, 640),    # more tiles than SMs: persistent loop + TMEM double buffering
    ],
)
def test_gemm_f32(ops, cuda_device, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(cuda_device).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(cuda_device)
    res = torch.randn(M, N, generator=g).to(cuda_device)
    out = ops.gemm(a, w, bias=bias, residual=res)
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    assert out.shape == (M, N)
    assert _rel(out, ref) < 2e-5  # fp32 accumulation of exact bf16 products vs fp64


@pytest.mark.parametrize("a_dt,w_dt", [(torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16)])
def test_gemm_operand_formats(ops, cuda_device, a_dt, w_dt):
    """The executor runs fp16 x fp16 where a norm bounds the activation's range and bf16 x bf16 elsewhere (a mixed
    bf16 x fp16 descriptor faults on sm_100a, which is why there is no such case here).  Exact products of the stored
    values, fp32 accumulation: the same tolerance for both formats."""
    g = torch.Generator().manual_seed(5)
    M, N, K = 384, 320, 1280
    a = torch.randn(M, K, generator=g).to(cuda_device).to(a_dt)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device).to(w_dt)
    bias = torch.randn(N, generator=g).to(cuda_device)
    out = ops.gemm_mixed(a, w, bias=bias)
    ref = a.double() @ w.double().t() + bias.double()
    assert _rel(out, ref) < 2e-5


def test_gemm_no_epilogue_and_bf16_out(ops, cuda_device):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(512, 256, generator=g).to(cuda_device).to(torch.bfloat16)
    w = (torch.randn(384, 256, generator=g) / 16).to(cuda_device).to(torch.bfloat16)
    ref = a.double() @ w.double().t()
    out = ops.gemm(a, w)
    assert _rel(out, ref) < 2e-5
    out16 = ops.gemm(a, w, out_mode=ops.OUT_BF16)
    assert out16.dtype == torch.bfloat16
    assert _rel(out16.float(), ref) < 4e-3  # one bf16 rounding of the output (2^-9 relative)


@pytest.mark.parametrize("M,N,K", [(1000, 320, 1280), (4096, 640, 640), (77, 64, 64)])
def test_gemm_residual_epilogue_bf16_and_f32(ops, cuda_device, M, N, K):
    """to_out / FF2 / proj_out (attention.py:311-326): GEMM + bias + fp32 residual through the coalescing epilogue
    (accumulator chunk transposed in shared memory), fp32 and bf16 outputs, ragged M."""
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda_device).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(cuda_device)
    res = (3.0 * torch.randn(M, N, generator=g)).to(cuda_device)
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    out = ops.gemm(a, w, bias=bias, residual=res)
    assert _rel(out, ref) < 2e-5
    out16 = ops.gemm(a, w, bias=bias, residual=res, out_mode=ops.OUT_BF16)
    assert out16.dtype == torch.bfloat16
    assert _rel(out16.float(), ref) < 4e-3
    out_nb = ops.gemm(a, w, residual=res)
    assert _rel(out_nb, ref - bias.double()) < 2e-5


def test_gemm_geglu(ops, cuda_device):
    # attention.py:68-75: x, gate = proj(x).chunk(2); x * gelu(gate)   (exact erf GELU)
    g = torch.Generator().manual_seed(2)
    M, C = 300, 128
    inner = 4 * C
    a = torch.randn(M, C, generator=g).to(cuda_device).to(torch.bfloat16)
    w = (torch.randn(2 * inner, C, generator=g) / math.sqrt(C)).to(cuda_device)
    b = torch.randn(2 * inner, generator=g).to(cuda_device)
    # interleave rows in blocks of 32: [x32 | gate32]
    wx, wg = w[:inner].reshape(-1, 32, C), w[inner:].reshape(-1, 32, C)
    wp = torch.stack([wx, wg], dim=1).reshape(2 * inner, C).to(torch.bfloat16).contiguous()
    bp = torch.stack([b[:inner].reshape(-1, 32), b[inner:].reshape(-1, 32)], dim=1).reshape(-1).contiguous()
    out = ops.gemm(a, wp, bias=bp, out_mode=ops.OUT_GEGLU)
    u = a.double() @ w.to(torch.bfloat16).double().t() + b.double()
    ref = u[:, :inner] * F.gelu(u[:, inner:])
    assert out.shape == (M, inner)
    assert _rel(out.float(), ref) < 4e-3


def test_gemm_tile_configurations_agree(ops, cuda_device, monkeypatch):
    """Every tile configuration (128- / 256-row CTA tiles, CTA pairs, several BN) and the opt-in
    autotuner compute the same GEMM: identical k order per output element, so bit-identical results."""
    g = torch.Generator().manual_seed(5)
    M, N, K = 2176, 960, 320  # 17 row tiles (odd: the pair kernel's last tile is half empty)
    a = torch.randn(M, K, generator=g).to(cuda_device).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(cuda_device)
    res = torch.randn(M, N, generator=g).to(cuda_device)
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    base = ops.gemm(a, w, bias=bias, residual=res)
    assert _rel(base, ref) < 2e-5
    for force in ("1,192", "2,96", "1,64", "2,64", "1,192,1", "1,64,1", "2,160"):
        monkeypatch.setenv("CAP4D_GEMM_FORCE", force)
        out = ops.gemm(a, w, bias=bias, residual=res)
        assert torch.equal(out, base), force
    monkeypatch.delenv("CAP4D_GEMM_FORCE")
    monkeypatch.setenv("CAP4D_GEMM_AUTOTUNE", "1")
    tuned = ops.gemm(a, w, bias=bias, residual=res)   # times the candidates, caches the winner
    tuned2 = ops.gemm(a, w, bias=bias, residual=res)  # cache hit
    assert torch.equal(tuned, base) and torch.equal(tuned2, base)


# ---------------------------------------------------------------------------------------------
# implicit-GEMM 3x3 convolution
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "n_img,H,W,Cin,Cout",
    [
        (2, 64, 64, 64, 64),     # W = 64: tile = 2 rows
        (3, 32, 32, 128, 96),    # W = 32: tile = 4 rows
        (4, 16, 16, 192, 128),   # tile = 8 rows
        (8, 8, 8, 256, 160),     # tile = 2 whole images
        (5, 4, 4, 64, 64),       # tile = 8 images, image tail -> OOB zero fill
        (8, 2, 2, 64, 32),       # lowest level of the tiny test model
        (16, 8, 8, 1280, 1280),  # production level-3 shape
        (1, 8, 256, 64, 64),     # image row wider than a tile (VAE decoder at 256 / 512 pixels): 2 tiles per row
        (2, 4, 512, 128, 32),    # 4 tiles per row
    ],
)
def test_conv3x3_stride1(ops, cuda_device, n_img, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(n_img * 100 + H)
    x = torch.randn(n_img, Cin, H, W, generator=g).to(cuda_device).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)).to(cuda_device)
    bias = torch.randn(Cout, generator=g).to(cuda_device)
    emb = torch.randn(n_img, Cout, generator=g).to(cuda_device)
    res = torch.randn(n_img * H * W, Cout, generator=g).to(cuda_device)
    wp = ops.pack_conv_weight(w)
    a = x.permute(0, 2, 3, 1).contiguous()
    out = ops.conv3x3(a, wp, n_img, H, W, stride=1, bias=bias, rowbias=emb, residual=res)
    ref = F.conv2d(x.double(), w.to(torch.bfloat16).double(), bias.double(), padding=1) + emb.double()[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).reshape(n_img * H * W, Cout) + res.double()
    assert _rel(out, ref) < 2e-5


@pytest.mark.parametrize("n_img,H,W,C", [(2, 64, 64, 64), (4, 16, 16, 128), (8, 4, 4, 64), (16, 16, 16, 640)])
def test_conv3x3_stride2(ops, cuda_device, n_img, H, W, C):
    # Downsample (openaimodel.py:150-153): conv3x3, stride 2, pad 1; H, W = INPUT size
    g = torch.Generator().manual_seed(H)
    x = torch.randn(n_img, C, H, W, generator=g).to(cuda_device).to(torch.bfloat16)
    w = (torch.randn(C, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(cuda_device)
    bias = torch.randn(C, generator=g).to(cuda_device)
    planes = ops.parity_planes(x.permute(0, 2, 3, 1).contiguous())
    out = ops.conv3x3(planes, ops.pack_conv_weight(w), n_img, H // 2, W // 2, stride=2, bias=bias)
    ref = F.conv2d(x.double(), w.to(torch.bfloat16).double(), bias.double(), stride=2, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, C)
    assert _rel(out, ref) < 2e-5


@pytest.mark.parametrize("n_img,H,W,C", [(2, 32, 32, 64), (4, 8, 8, 128), (8, 2, 2, 64), (16, 16, 16, 1280)])
def test_upsample_conv3x3(ops, cuda_device, n_img, H, W, C):
    # Upsample (openaimodel.py:111-119): nearest 2x then conv3x3; here four phase convs with pre-summed taps.
    # The summed taps are rounded to bf16 once, so the comparison uses the same summed-then-rounded weights.
    g = torch.Generator().manual_seed(H + C)
    x = torch.randn(n_img, C, H, W, generator=g).to(cuda_device).to(torch.bfloat16)
    w = (torch.randn(C, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(cuda_device)
    bias = torch.randn(C, generator=g).to(cuda_device)
    out = ops.upsample_conv3x3(x.permute(0, 2, 3, 1).contiguous(), w, bias=bias)
    up = F.interpolate(x.double(), scale_factor=2, mode="nearest")
    exact = F.conv2d(up, w.double(), bias.double(), padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    # exact-weight reference: only the bf16 rounding of the (summed) weights separates the two
    assert _rel(out, exact) < 6e-3
    assert float((out.double() - exact).norm() / exact.norm()) < 3e-3


# ---------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------
def _attn_ref(qkv, C, L):
    M = qkv.shape[0]
    heads = C // 64
    q, k, v = qkv.double().split(C, dim=1)

    def sp(t):
        return t.reshape(M // L, L, heads, 64).permute(0, 2, 1, 3)

    s = sp(q) @ sp(k).transpose(-1, -2) * 0.125
    o = s.softmax(-1) @ sp(v)
    return o.permute(0, 2, 1, 3).reshape(M, C)


@pytest.mark.parametrize(
    "n_seq,L,C",
    [
        (1, 128, 64),     # single tile
        (2, 256, 128),    # two kv tiles, two heads
        (1, 1024, 64),    # kv ring wrap-around (KS = VS = 3), S double buffering
        (3, 64, 64),      # L < tile: key masking + neighbouring-sequence rows in the box
        (2, 200, 64),     # ragged: q tail + kv tail
        (8, 16, 256),     # tiny sequences (lowest level of the test model)
        (2, 2048, 1280),  # production level-2 "3d" shape (V=8 x 16x16), 20 heads
    ],
)
def test_attention(ops, cuda_device, n_seq, L, C):
    g = torch.Generator().manual_seed(L + C)
    qkv = torch.randn(n_seq * L, 3 * C, generator=g).to(cuda_device).to(torch.bfloat16)
    out = ops.attention(qkv, C, L)
    ref = _attn_ref(qkv, C, L)
    # P and the output are rounded to bf16 (2^-9); everything else is fp32
    assert _rel(out.float(), ref) < 1e-2
    assert float((out.double() - ref).norm() / ref.norm()) < 5e-3


@pytest.mark.parametrize("L,slope,spike", [(1024, 1.5, 0.0), (1024, 0.0, 20.0), (1000, 0.4, 25.0), (200, 0.0, 30.0)])
def test_attention_logits_that_outgrow_the_first_tile(ops, cuda_device, L, slope, spike):
    # the kernel takes the exact row max at KV tile 0 only and follows later growth through the row sums
    # (attn_tc.cu, rq_regrow): logits that climb by ~17 octaves per tile (slope) exercise the power-of-two
    # shift, a late key ~100+ nats above everything before it (spike) the recomputation from global memory.
    # legacy_attention's softmax (attention.py:125) is exact for both.
    g = torch.Generator().manual_seed(L + int(10 * slope) + int(spike))
    C = 128
    base = torch.randn(1, C, generator=g)
    base = base / base.reshape(2, 64).norm(dim=1).repeat_interleave(64) * 8.0   # |base_h|^2 = 64 per head
    q = base + 0.1 * torch.randn(L, C, generator=g)
    tile = (torch.arange(L) // 128).float().unsqueeze(1)
    k = torch.randn(L, C, generator=g) * 0.5 + slope * tile * base
    if spike:
        k[L - 5] = spike * base[0]
        k[L // 2 + 3] = 0.5 * spike * base[0]
    v = torch.randn(L, C, generator=g)
    qkv = torch.cat([q, k, v], dim=1).to(cuda_device).to(torch.bfloat16)
    out = ops.attention(qkv, C, L)
    ref = _attn_ref(qkv, C, L)
    assert torch.isfinite(out.float()).all()
    assert _rel(out.float(), ref) < 1e-2
    assert float((out.double() - ref).norm() / ref.norm()) < 5e-3


def test_attention_is_key_permutation_invariant(ops, cuda_device):
    # the "3d" rearrange '(b t) n (h d) -> (b h) (n t) d' only permutes tokens inside a sequence
    g = torch.Generator().manual_seed(5)
    L, C = 512, 64
    qkv = torch.randn(L, 3 * C, generator=g).to(cuda_device).to(torch.bfloat16)
    perm = torch.randperm(L, generator=g).to(cuda_device)
    out = ops.attention(qkv, C, L)
    out_p = ops.attention(qkv[perm].contiguous(), C, L)
    assert _rel(out_p.float(), out[perm].float()) < 1e-2


# ---------------------------------------------------------------------------------------------
# norms
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "n_img,hw,C1,C2,silu,eps",
    [
        (2, 256, 64, 0, True, 1e-5),
        (4, 64, 320, 0, False, 1e-6),       # transformer Normalize, cpg = 10 (quads straddle groups)
        (3, 1024, 640, 320, True, 1e-5),    # concat seam inside a group (960 / 32 = 30 per group, seam at 640)
        (16, 64, 1280, 1280, True, 1e-5),   # 2560 channels
        (8, 4, 128, 64, True, 1e-5),        # tiny hw
        (2, 4096, 320, 0, True, 1e-5),      # production level-0 row count
    ],
)
def test_groupnorm(ops, cuda_device, n_img, hw, C1, C2, silu, eps):
    g = torch.Generator().manual_seed(C1 + C2 + hw)
    C = C1 + C2
    x = (torch.randn(n_img * hw, C, generator=g) * 2 + torch.randn(C, generator=g)).to(cuda_device)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(cuda_device)
    beta = (0.2 * torch.randn(C, generator=g)).to(cuda_device)
    x1 = x[:, :C1].contiguous()
    x2 = x[:, C1:].contiguous() if C2 else None
    out, raw = ops.groupnorm(x1, x2, n_img, hw, gamma, beta, eps, silu, want_raw=True)
    xr = x.reshape(n_img, hw, C).permute(0, 2, 1).double()
    ref = F.group_norm(xr, 32, gamma.double(), beta.double(), eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(n_img * hw, C)
    assert _rel(out.float(), ref) < 5e-3       # bf16 output rounding
    assert torch.equal(raw, x.to(torch.bfloat16))


@pytest.mark.parametrize("M,C", [(1000, 64), (4096, 320), (512, 1280), (7, 640)])
def test_layernorm(ops, cuda_device, M, C):
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(M, C, generator=g) * 3 + 1).to(cuda_device)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(cuda_device)
    beta = (0.2 * torch.randn(C, generator=g)).to(cuda_device)
    out = ops.layernorm(x, gamma, beta, 1e-5)
    ref = F.layer_norm(x.double(), (C,), gamma.double(), beta.double(), 1e-5)
    assert _rel(out.float(), ref) < 5e-3


# ---------------------------------------------------------------------------------------------
# CFG + DDIM update
# ---------------------------------------------------------------------------------------------
def test_cfg_ddim_update_bit_exact(ops, cuda_device):
    # sampler.py:205-231 in eager fp32 ops is the reference arithmetic; the fused kernel must match
    # bit for bit (separately rounded multiplies/adds, no FMA contraction)
    g = torch.Generator().manual_seed(9)
    n_groups, V, R, chw, n_gen = 3, 8, 1, 4 * 16 * 16, 40
    lat = torch.randn(n_gen, chw, generator=g).to(cuda_device)
    eps = torch.randn(2 * n_groups, V, chw, generator=g).to(cuda_device)
    idx = torch.randperm(n_gen, generator=g)[: n_groups * (V - R)].reshape(n_groups, V - R).to(cuda_device)
    cfg, xf, ef = 2.0, 1.0123457, -0.0456789
    ref = lat.clone()
    eu, ec = eps[:n_groups], eps[n_groups:]
    e = (eu + cfg * (ec - eu))[:, R:]
    e_all = torch.zeros_like(ref)
    e_all[idx.reshape(-1)] += e.reshape(-1, chw)
    upd = ref * torch.tensor(xf).float().to(cuda_device) + e_all * torch.tensor(ef).float().to(cuda_device)
    ref[idx.reshape(-1)] = upd[idx.reshape(-1)]
    ops.cfg_ddim_update(lat, eps, idx, n_groups, V, R, cfg, xf, ef)
    assert torch.equal(lat, ref)
