"""-m gpu: every C-ABI kernel against a plain PyTorch fp32 statement of the same op (tolerances written per test)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as ge
    ge.build()
    from dig_b200 import ops as o
    o.load()
    assert o.load().dig_sm() == 100, "dig_b200 is built for sm_100a (B200)"
    return o


def rnd(*shape, scale=1.0, dtype=torch.float32):
    return (torch.randn(*shape, device="cuda") * scale).to(dtype)


@pytest.mark.parametrize("amn,bmn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 384), (360, 192, 384), (1024, 1152, 384), (200, 48, 192), (512, 384, 48)])
def test_gemm_operand_majors_and_tails(ops, amn, bmn, M, N, K):
    torch.manual_seed(1)
    if amn and M % 8:
        pytest.skip("MN-major A needs lda % 8 == 0")
    a = rnd(K, M, dtype=torch.bfloat16) if amn else rnd(M, K, dtype=torch.bfloat16)
    b = rnd(K, N, dtype=torch.bfloat16) if bmn else rnd(N, K, dtype=torch.bfloat16)
    ref = (a.float().t() if amn else a.float()) @ (b.float() if bmn else b.float().t())
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, out, a_mn_major=amn, b_mn_major=bmn)
    # fp32 accumulation of exact bf16 products: only summation order differs
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-3 * K ** 0.5)
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, outb, a_mn_major=amn, b_mn_major=bmn)
    assert torch.allclose(outb.float(), ref, rtol=8e-3, atol=8e-3 * K ** 0.5)


def test_gemm_split_k_accumulates(ops):
    torch.manual_seed(2)
    K, M, N = 8192, 384, 1152
    a, b = rnd(K, M, dtype=torch.bfloat16), rnd(K, N, dtype=torch.bfloat16)
    base = rnd(M, N)
    out = base.clone()
    ops.gemm(a, b, out, a_mn_major=True, b_mn_major=True, split_k=7)
    ref = base + a.float().t() @ b.float()
    assert torch.allclose(out, ref, rtol=1e-4, atol=5e-2)


@pytest.mark.parametrize("M,N", [(384, 1536), (1536, 384), (1152, 384), (384, 384), (48, 192), (192, 384)])
def test_gemm_auto_split_k_weight_gradient_shapes(ops, M, N):
    """split_k=-1: the library picks tile shape and K slices (2-CTA 256-wide pair tiles with ragged, clipped last tiles)."""
    torch.manual_seed(7)
    K = 4096 + 64
    a, b = rnd(K, M, dtype=torch.bfloat16), rnd(K, N, dtype=torch.bfloat16)
    base = rnd(M, N)
    out = base.clone()
    ops.gemm(a, b, out, a_mn_major=True, b_mn_major=True, split_k=-1)
    ref = base + a.float().t() @ b.float()
    assert torch.allclose(out, ref, rtol=1e-4, atol=5e-2)


def test_gemm_fused_epilogues(ops):
    torch.manual_seed(3)
    M, N, K = 640, 384, 384
    a, w = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=0.05, dtype=torch.bfloat16)
    bias, res = rnd(N), rnd(M, N)
    acc = a.float() @ w.float().t()
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, out, bias=bias, residual=res)
    assert torch.allclose(out, acc + bias + res, atol=2e-4)
    # GELU forward: bf16 post-activation + bf16 pre-activation (F:54-55)
    post, pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, post, bias=bias, epilogue=ops.EPI_GELU, aux=pre)
    assert torch.allclose(pre.float(), acc + bias, atol=2e-2)
    assert torch.allclose(post.float(), torch.nn.functional.gelu(acc + bias), atol=2e-2)
    # GELU backward with fused column sums
    wt = w.t().contiguous()   # [K, N] MN-major B
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    dh, cs = torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.zeros(N, device="cuda")
    ops.gemm(a, wt, dh, b_mn_major=True, epilogue=ops.EPI_GELU_BWD, aux=pre, colsum=cs)
    refb = acc * x.grad
    assert torch.allclose(dh.float(), refb, atol=3e-2, rtol=1e-2)
    # the fused bias gradient sums the bf16-rounded dh tile that is staged for the TMA store (as autocast's autograd would)
    assert torch.allclose(cs, dh.float().sum(0), atol=2e-2, rtol=1e-4)
    assert torch.allclose(cs, refb.sum(0), atol=0.15, rtol=2e-3)
    # ROWDOT: bf16 output + per-64-column-group dot product with aux (D of the attention backward)
    o3, rd = torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.zeros(M, N // 64, device="cuda")
    ops.gemm(a, wt, o3, b_mn_major=True, epilogue=ops.EPI_ROWDOT, aux=pre, rowdot=rd)
    assert torch.allclose(o3.float(), acc, atol=3e-2, rtol=1e-2)
    assert torch.allclose(rd, (acc * pre.float()).view(M, N // 64, 64).sum(-1), atol=5e-2, rtol=1e-2)
    # ReLU mask
    act = torch.relu(rnd(M, N)).to(torch.bfloat16)
    o2 = torch.empty(M, N, device="cuda")
    ops.gemm(a, wt, o2, b_mn_major=True, epilogue=ops.EPI_RELU_MASK, aux=act)
    assert torch.allclose(o2, acc * (act.float() > 0), atol=2e-4)
    # patch-embed epilogue: mask-token rows + position table (V:95-99)
    mask = (torch.rand(M, device="cuda") < 0.5).to(torch.uint8)
    tok, pos = rnd(N), rnd(128, N)
    ops.gemm(a, w, out, bias=bias, residual=pos, res_row_mod=128, row_mask=mask, row_mask_value=tok)
    ref = torch.where(mask.bool()[:, None], tok.expand(M, N), acc + bias) + pos.repeat(M // 128, 1)
    assert torch.allclose(out, ref, atol=2e-4)


def q8_codes(pre):
    """The 8-bit pre-activation code of include/dig_b200.h (dig_gemm_t.aux_q8), restated with torch ops."""
    return torch.round((pre.clamp(-4.0, 4.0) + 4.0) * (255.0 / 8.0)).to(torch.uint8)


def q8_gelu_grad(codes):
    x = (codes.double() * (8.0 / 255.0) - 4.0).requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    return x.grad.float()


@pytest.mark.parametrize("M,N,K", [(640, 384, 384), (4096, 1536, 384), (200, 192, 128)])
def test_gemm_gelu_epilogues_with_8bit_preactivation_codes(ops, M, N, K):
    """DIG_EPI_GELU / DIG_EPI_GELU_BWD with aux_q8 (1-CTA kernel, TMA and generic epilogues): codes match the restated quantiser up
    to one level at rounding ties, the backward equals dy.W x gelu'(decoded level), and it stays within 1.3e-2 of the exact gelu'."""
    torch.manual_seed(5)
    a, w = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=2.0 / K ** 0.5, dtype=torch.bfloat16)
    bias = rnd(N)
    acc = a.float() @ w.float().t() + bias
    post, codes = torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.zeros(M, N, device="cuda", dtype=torch.uint8)
    ops.gemm(a, w, post, bias=bias, epilogue=ops.EPI_GELU, aux=codes)
    assert torch.allclose(post.float(), torch.nn.functional.gelu(acc), atol=2e-2, rtol=8e-3)
    diff = (codes.int() - q8_codes(acc).int()).abs()
    assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 2e-3      # fp32 accumulation order moves a value across a tie
    assert acc.abs().max() > 4.0 and int(codes.min()) == 0 and int(codes.max()) == 255   # the clamp is exercised
    wt = w.t().contiguous()
    dh, cs = torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.zeros(N, device="cuda")
    ops.gemm(a, wt, dh, b_mn_major=True, epilogue=ops.EPI_GELU_BWD, aux=codes, colsum=cs)
    plain = a.float() @ w.float().t()
    ref = plain * q8_gelu_grad(codes)
    assert torch.allclose(dh.float(), ref, atol=3e-2, rtol=1e-2)
    assert torch.allclose(cs, dh.float().sum(0), atol=5e-2, rtol=1e-3)
    x = acc.clone().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    assert float((q8_gelu_grad(codes) - x.grad).abs().max()) < 1.3e-2
    assert float((q8_gelu_grad(codes) - x.grad).pow(2).mean().sqrt()) < 5e-3


def test_gemm_unbuilt_combination_is_an_error(ops):
    a = rnd(128, 64, dtype=torch.bfloat16)
    with pytest.raises(ops.DigError):
        ops.gemm(a, a, torch.empty(128, 128, device="cuda"), epilogue=ops.EPI_GELU, aux=torch.empty(128, 128, device="cuda", dtype=torch.bfloat16))


@pytest.mark.parametrize("heads,S", [(6, 3), (3, 2), (8, 1)])
def test_attention_forward_backward(ops, heads, S):
    torch.manual_seed(4)
    d, scale = heads * 64, 64 ** -0.5
    qkv = rnd(S * 256, 3 * d, scale=1.5, dtype=torch.bfloat16)
    x = qkv.float().requires_grad_(True)
    q, k, v = x.view(S, 256, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q * scale) @ k.transpose(-1, -2)
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(S * 256, d)
    lse_ref = torch.logsumexp(s, -1)
    for smem_p in (False, True):
        out = torch.empty(S * 256, d, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(S, heads, 256, device="cuda")
        ops.attention_fwd(qkv, out, lse, heads, scale, p_in_smem=smem_p)
        assert torch.allclose(out.float(), ref, atol=3e-2), smem_p       # bf16 P and bf16 output
        assert torch.allclose(lse, lse_ref, atol=1e-4)
    dout = rnd(S * 256, d, dtype=torch.bfloat16)
    ref.backward(dout.float())
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(qkv, out, dout, lse, dqkv, heads, scale)
    assert torch.allclose(dqkv.float(), x.grad, atol=6e-2, rtol=2e-2)   # bf16 P / dS operands
    # persistent kernel with D = rowsum(dO o O) supplied (as the output-projection dgrad epilogue emits it)
    dsum = (dout.float() * out.float()).view(S * 256, heads, 64).sum(-1).contiguous()
    dqkv2 = torch.zeros_like(qkv)
    ops.attention_bwd_d(qkv, dout, lse, dsum, dqkv2, heads, scale)
    assert torch.allclose(dqkv2.float(), x.grad, atol=6e-2, rtol=2e-2)


@pytest.mark.parametrize("d,gelu", [(384, 0), (512, 0), (192, 1), (192, 0)])
def test_layernorm_forward_backward(ops, d, gelu):
    torch.manual_seed(5)
    rows = 1000
    x = rnd(rows, d, scale=2.0).requires_grad_(True)
    g, b = (1 + 0.1 * rnd(d)).requires_grad_(True), (0.1 * rnd(d)).requires_grad_(True)
    ref = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-6)
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    y = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.call("dig_layernorm_fwd", x.detach(), g.detach(), b.detach(), y, mean, rstd, rows, d, 1e-6, gelu)
    assert torch.allclose(y.float(), ref, atol=2e-2)
    dy = rnd(rows, d, dtype=torch.bfloat16)
    dres = rnd(rows, d)
    ref.backward(dy.float())
    dx, dxb = torch.empty(rows, d, device="cuda"), torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    dg, db, dsum = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    ops.call("dig_layernorm_bwd", dy, x.detach(), mean, rstd, g.detach(), b.detach(), dres, dx, dxb, dg, db, dsum, rows, d, gelu)
    assert torch.allclose(dx, x.grad + dres, atol=1e-4, rtol=1e-4)
    assert torch.allclose(dxb.float(), dx, atol=3e-2, rtol=1e-2)
    assert torch.allclose(dg, g.grad, atol=1e-3, rtol=1e-4) and torch.allclose(db, b.grad, atol=1e-3, rtol=1e-4)
    assert torch.allclose(dsum, dx.sum(0), atol=1e-3, rtol=1e-4)
    # in-place residual-gradient update (dx_f32 aliases dres), as the encoder backward uses it
    d2 = dres.clone()
    ops.call("dig_layernorm_bwd", dy, x.detach(), mean, rstd, g.detach(), b.detach(), d2, d2, None, dg, db, None, rows, d, gelu)
    assert torch.allclose(d2, dx, atol=1e-5)


@pytest.mark.parametrize("d", [384, 512, 192])
def test_layernorm_backward_with_bf16_residual_gradient_stream(ops, d):
    """dig_layernorm_bwd_bf16res: dx_bf16 = bf16(dres_bf16 + dLN(dy)); out of place and in place (the single-stream schedule)."""
    torch.manual_seed(6)
    rows = 777
    x = rnd(rows, d, scale=2.0).requires_grad_(True)
    g = (1 + 0.1 * rnd(d)).requires_grad_(True)
    b = (0.1 * rnd(d)).requires_grad_(True)
    y = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.call("dig_layernorm_fwd", x.detach(), g.detach(), b.detach(), y, mean, rstd, rows, d, 1e-6, 0)
    dy, dres = rnd(rows, d, dtype=torch.bfloat16), rnd(rows, d, dtype=torch.bfloat16)
    torch.nn.functional.layer_norm(x, (d,), g, b, 1e-6).backward(dy.float())
    ref = x.grad + dres.float()
    dxb = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    dg, db, dsum = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    ops.call("dig_layernorm_bwd_bf16res", dy, x.detach(), mean, rstd, g.detach(), dres, dxb, dg, db, dsum, rows, d)
    assert torch.allclose(dxb.float(), ref, atol=3e-2, rtol=1e-2)
    assert torch.allclose(dg, g.grad, atol=1e-3, rtol=1e-4) and torch.allclose(db, b.grad, atol=1e-3, rtol=1e-4)
    assert torch.allclose(dsum, ref.sum(0), atol=2e-3, rtol=1e-4)          # column sums of the fp32 values before rounding
    inpl = dres.clone()
    ops.call("dig_layernorm_bwd_bf16res", dy, x.detach(), mean, rstd, g.detach(), inpl, inpl, dg, db, None, rows, d)
    assert torch.equal(inpl, dxb)
    # masked-row zeroing of a bf16 stream
    mask = (torch.rand(rows, device="cuda") < 0.5).to(torch.uint8)
    gz = torch.empty_like(dxb)
    ops.call("dig_zero_masked_rows_bf16", dxb, mask, gz, rows, d)
    assert torch.equal(gz, torch.where(mask.bool()[:, None], torch.zeros_like(dxb), dxb))


def test_batchnorm_forward_backward(ops):
    torch.manual_seed(6)
    rows, C = 517, 512
    x = rnd(rows, C, scale=3.0).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(C).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    ref = torch.relu(bn(x))
    stats = torch.zeros(2 * C, device="cuda")
    ops.call("dig_colsum", x.detach(), 1, C, stats, stats[C:], rows, C)
    y, yf = torch.empty(rows, C, device="cuda", dtype=torch.bfloat16), torch.empty(rows, C, device="cuda")
    ops.call("dig_bn_apply", x.detach(), stats, float(rows), bn.weight, bn.bias, 1, bn.eps, y, yf, rows, C)
    assert torch.allclose(yf, ref, atol=1e-4)
    rm, rv, nbt = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    ops.call("dig_bn_running", stats, float(rows), 0.1, rm, rv, nbt, C)
    assert torch.allclose(rm, bn.running_mean, atol=1e-5) and torch.allclose(rv, bn.running_var, atol=1e-4) and int(nbt) == 1
    dout = rnd(rows, C)
    ref.backward(dout)
    dy = dout * (ref.detach() > 0)
    bst = torch.zeros(2 * C, device="cuda")
    ops.call("dig_bn_bwd_stats", dy, x.detach(), stats, float(rows), bn.eps, bst, rows, C)
    assert torch.allclose(bst[:C], bn.bias.grad, atol=1e-3, rtol=1e-4) and torch.allclose(bst[C:], bn.weight.grad, atol=1e-3, rtol=1e-4)
    dx = torch.empty(rows, C, device="cuda")
    ops.call("dig_bn_bwd_apply", dy, x.detach(), stats, bst, float(rows), bn.weight, bn.eps, None, dx, rows, C)
    assert torch.allclose(dx, x.grad, atol=1e-4, rtol=1e-3)


def test_row_kernels(ops):
    torch.manual_seed(7)
    S, d, nw = 6, 384, 4
    img = torch.rand(S, 3, 32, 128, device="cuda") * 2 - 1
    a0 = torch.empty(S * 256, 48, device="cuda", dtype=torch.bfloat16)
    ops.call("dig_im2col_patch4", img, a0, S)
    ref = torch.nn.functional.unfold(img, 4, stride=4).transpose(1, 2).reshape(S * 256, 48)   # (c, kh, kw) column order
    assert torch.equal(a0, ref.to(torch.bfloat16))
    x0, x1 = rnd(3 * 256, d), rnd(3 * 256, d)
    pooled = torch.empty(S * nw, d, device="cuda", dtype=torch.bfloat16)
    ops.call("dig_pool_fwd", x0, x1, 3, pooled, S, d, nw)
    xc = torch.cat([x0, x1]).view(S, 8, 32, d).permute(0, 3, 1, 2)
    refp = torch.nn.functional.adaptive_avg_pool2d(xc, (1, nw)).permute(0, 2, 3, 1).reshape(S * nw, d)
    assert torch.allclose(pooled.float(), refp, atol=1e-2)
    dp = rnd(S * nw, d)
    d0, d1 = torch.empty_like(x0), torch.empty_like(x1)
    ops.call("dig_pool_bwd", dp, 3, d0, d1, S, d, nw)
    refd = dp.view(S, 1, nw, 1, d).expand(S, 8, nw, 8, d).reshape(S * 256, d) / 64
    assert torch.allclose(torch.cat([d0, d1]), refd, atol=1e-6)
    mask = torch.zeros(S, 256, dtype=torch.bool, device="cuda")
    for b_ in range(S):
        mask[b_, torch.randperm(256, device="cuda")[:179]] = True
    idx, err = torch.empty(S * 179, dtype=torch.int32, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.call("dig_mask_to_index", mask.to(torch.uint8), idx, err, S, 179)
    assert int(err) == 0 and torch.equal(idx.long(), mask.flatten().nonzero().flatten())
    ops.call("dig_mask_to_index", mask.to(torch.uint8), idx, err, S, 178)
    assert int(err) == 1                                               # ragged masks are flagged
    x = rnd(S * 256, d)
    ops.call("dig_mask_to_index", mask.to(torch.uint8), idx, err, S, 179)
    gat = torch.empty(S * 179, d, device="cuda", dtype=torch.bfloat16)
    ops.call("dig_gather_rows", x, idx, gat, S * 179, d)
    assert torch.equal(gat, x[mask.flatten()].to(torch.bfloat16))
    dst = rnd(S * 256, d)
    ref = dst.clone()
    src = rnd(S * 179, d)
    ref[mask.flatten()] += src
    ops.call("dig_scatter_add_rows", src, idx, dst, S * 179, d)
    assert torch.allclose(dst, ref)
    xb = rnd(4099, 1152, dtype=torch.bfloat16)
    cs = torch.zeros(1152, device="cuda")
    ops.call("dig_colsum", xb, 0, 1152, cs, None, 4099, 1152)
    assert torch.allclose(cs, xb.float().sum(0), atol=2e-2, rtol=1e-4)


def test_contrastive_and_pixel_losses(ops):
    torch.manual_seed(8)
    Q, Nk, C, T, off = 40, 120, 256, 0.2, 40
    q = rnd(Q, C).requires_grad_(True)
    k = torch.nn.functional.normalize(rnd(Nk, C), dim=1)
    qn_ref = torch.nn.functional.normalize(q, dim=1)
    logits = qn_ref @ k.t() / T
    labels = torch.arange(Q, device="cuda") + off
    loss_ref = torch.nn.functional.cross_entropy(logits, labels) * 2 * T
    loss_ref.backward()
    qn, inv = torch.empty(Q, C, device="cuda"), torch.empty(Q, device="cuda")
    ops.call("dig_l2norm_fwd", q.detach(), qn, inv, Q, C)
    lg = torch.empty(Q, Nk, device="cuda")
    ops.call("dig_sgemm_f32", qn, k, lg, Q, Nk, C, 1, 1.0 / T)
    assert torch.allclose(lg, logits.detach(), atol=1e-5)
    res = torch.zeros(4, device="cuda")
    ops.call("dig_infonce_rows", lg, Q, Nk, off, T, res)
    assert float(res[0]) == pytest.approx(float(loss_ref), rel=1e-5)
    top = logits.detach().topk(5, 1).indices.eq(labels[:, None])
    assert float(res[1]) == pytest.approx(float(top[:, :1].sum()) * 100 / Q, abs=1e-3)
    assert float(res[2]) == pytest.approx(float(top.sum()) * 100 / Q, abs=1e-3)
    dqn, dq = torch.empty(Q, C, device="cuda"), torch.empty(Q, C, device="cuda")
    ops.call("dig_sgemm_f32", lg, k, dqn, Q, C, Nk, 0, 1.0)
    ops.call("dig_l2norm_bwd", dqn, qn, inv, torch.ones(1, device="cuda"), dq, Q, C)
    assert torch.allclose(dq, q.grad, atol=1e-6, rtol=1e-4)
    # masked-pixel MSE against the engine's einops statement (E:85-111, E:141)
    from oracle import restatement as R
    B = 5
    img, _, mask = R.synthetic_batch(B, seed=3)
    mk = mask.clone()
    mk[:, 1] = False
    tgt = R.build_targets(img, mk)[0].cuda()
    pred = rnd(B, 179, 48).requires_grad_(True)
    ref = torch.nn.functional.mse_loss(pred, tgt)
    ref.backward()
    from dig_b200.engine import masked_pixel_mse
    p2 = pred.detach().clone().requires_grad_(True)
    out = masked_pixel_mse(p2, img.cuda(), mk[:, 0].cuda())
    (out * 3.0).backward()
    assert float(out) == pytest.approx(float(ref), rel=1e-5)
    assert torch.allclose(p2.grad, 3.0 * pred.grad, atol=1e-8, rtol=1e-4)
    # the `normlize_target=True` branch (E:89-94): per-patch, per-colour standardised targets
    tgt_n = R.build_targets(img, mk, normalize_target=True)[0].cuda()
    pred_n = pred.detach().clone().requires_grad_(True)
    ref_n = torch.nn.functional.mse_loss(pred_n, tgt_n)
    ref_n.backward()
    p3 = pred.detach().clone().requires_grad_(True)
    out_n = masked_pixel_mse(p3, img.cuda(), mk[:, 0].cuda(), normalize_target=True)
    out_n.backward()
    assert float(out_n) == pytest.approx(float(ref_n), rel=2e-5) and abs(float(out_n) - float(out)) > 1e-3
    assert torch.allclose(p3.grad, pred_n.grad, atol=1e-7, rtol=1e-3)


def test_multi_tensor_kernels(ops):
    torch.manual_seed(9)
    from dig_b200.optim import FusedAdamW
    from dig_b200.pretrain_step import MtTable
    from dig_b200.utils import get_grad_norm_
    shapes = [(384, 384), (1536,), (3, 5, 7), (20000,)]
    ps = [torch.nn.Parameter(rnd(*s)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    for p, q in zip(ps, qs):
        p.grad = rnd(*p.shape)
        q.grad = p.grad.clone()
    assert float(get_grad_norm_(ps)) == pytest.approx(float(torch.norm(torch.stack([p.grad.norm() for p in ps]))), rel=1e-5)
    a = FusedAdamW([{"params": ps[:2], "weight_decay": 0.05, "lr_scale": 1.0}, {"params": ps[2:], "weight_decay": 0.0, "lr_scale": 1.0}], lr=1e-2)
    b = torch.optim.AdamW([{"params": qs[:2], "weight_decay": 0.05}, {"params": qs[2:], "weight_decay": 0.0}], lr=1e-2)
    for _ in range(3):
        a.step()
        b.step()
    for p, q in zip(ps, qs):
        assert torch.allclose(p, q, atol=1e-6, rtol=1e-5)
    online = [rnd(1000), rnd(33, 9)]
    target = [rnd(1000), rnd(33, 9)]
    shadow = [torch.empty(1000, device="cuda", dtype=torch.bfloat16), None]
    ref = [t * 0.9 + o * 0.1 for o, t in zip(online, target)]
    tab = MtTable(torch.device("cuda"), online, target, shadow)
    lib = ops.load()
    rc = lib.dig_mt_ema(tab.ptrs[0].data_ptr(), tab.ptrs[1].data_ptr(), tab.ptrs[2].data_ptr(), tab.numel.data_ptr(), tab.blk_tensor.data_ptr(),
                        tab.blk_chunk.data_ptr(), tab.num_blocks, 0.9, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert torch.allclose(target[0], ref[0], atol=1e-6) and torch.allclose(target[1], ref[1], atol=1e-6)
    assert torch.equal(shadow[0], target[0].to(torch.bfloat16))
