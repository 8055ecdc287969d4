"""GPU parity tests of the C-ABI kernels against the CPU oracle (`oracle/oracle.py`).
Everything here calls libvqb200.so through ctypes (`vector_quantization_b200.ops`)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from vector_quantization_b200 import ops

pytestmark = pytest.mark.gpu

IDX_EPS = 1e-5  # accepted near-tie: oracle distance gap of a differing index, relative to max(1, d)


def _check_indices(d_oracle, q_oracle, q_test, eps=IDX_EPS, what='', squared=False):
    """squared=True: near-ties are judged on d^2.  torch.cdist's mm path computes sqrt(|x|^2 - 2 x.e + |e|^2), whose
    fp32 cancellation error is ~1e-7 (|x|^2 + |e|^2) on d^2: when the nearest codes are much closer than the vector
    norms (e.g. D = 1 with thousands of codes) the reference's own d is only meaningful to that resolution."""
    if squared:
        d_oracle = d_oracle * d_oracle
    rows, gap = O.index_mismatch_report(d_oracle, q_oracle, q_test)
    if rows.numel():
        scale = d_oracle[rows, q_oracle[rows]].abs().clamp_min(1.0)
        bad = gap > eps * scale
        assert not bad.any(), (f'{what}: {int(bad.sum())} index mismatches beyond the near-tie epsilon '
                               f'(max gap {float(gap.max()):.3e}, of {rows.numel()} differing rows)')
    return rows.numel()


def _assign(x, E, metric, dev, backend, planes_x=None, planes_e=None, normalize_e=None):
    """row arg-min of the reference distance through vqb_pack_rows + vqb_assign."""
    cos = metric == 'Cosine'
    xe = x.to(dev)
    Ee = E.to(dev)
    a = ops.pack_rows(xe, normalize=False, planes=planes_x)
    b = ops.pack_rows(Ee, normalize=cos if normalize_e is None else normalize_e, planes=planes_e,
                      want_half_sqnorm=not cos)
    keys = ops.new_keys(x.shape[0], dev)
    ops.assign(a, b, keys, l2=not cos, backend=backend)
    return ops.unpack_keys(keys, want_score=True)


# --------------------------------------------------------------------------------------------


@pytest.mark.parametrize('rows,D,normalize,planes,dtype', [
    (1000, 32, False, 3, torch.float32), (513, 8, True, 3, torch.float32), (300, 256, True, 2, torch.float32),
    (777, 32, False, 1, torch.bfloat16), (64, 5, False, 3, torch.float32), (130, 768, True, 3, torch.bfloat16)])
def test_pack_rows(dev, rows, D, normalize, planes, dtype):
    g = torch.Generator().manual_seed(1)
    src = torch.randn(rows, D, generator=g).to(dtype)
    wb = torch.empty(rows, D, dtype=torch.float32, device=dev)
    op = ops.pack_rows(src.to(dev), normalize=normalize, planes=planes, want_half_sqnorm=True, writeback=wb)
    ref = F.normalize(src.float()) if normalize else src.float()
    rows_pad, Dp = ops.operand_shape(rows, D)
    assert op.planes.shape == (planes, rows_pad, Dp)
    recon = op.planes.float().sum(0).cpu()
    tol = {1: 2 ** -8, 2: 2 ** -16, 3: 1e-7}[planes] if not (dtype == torch.bfloat16 and not normalize) else 0.0
    assert torch.allclose(recon[:rows, :D], ref, rtol=tol, atol=1e-7 if tol else 0.0)
    assert (recon[rows:] == 0).all() and (recon[:, D:] == 0).all()
    torch.testing.assert_close(wb.cpu(), ref, rtol=2e-6, atol=1e-7)
    h = op.half_sqnorm.cpu()
    torch.testing.assert_close(h[:rows], 0.5 * (ref * ref).sum(1), rtol=1e-5, atol=1e-7)
    assert torch.isinf(h[rows:]).all()
    if planes == 3:  # the three planes reproduce the fp32 value bit for bit
        assert torch.equal(recon[:rows, :D], wb.cpu())


@pytest.mark.parametrize('backend', [ops.BACKEND_SIMT, ops.BACKEND_TCGEN05], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('metric', ['L2', 'Cosine'])
@pytest.mark.parametrize('N,K,D', [(1000, 700, 32), (2048, 1024, 8), (384, 512, 256), (300, 333, 64), (257, 4100, 20),
                                   (128, 256, 768), (1, 1, 1), (3, 2, 5), (130, 7, 16), (5, 300, 2)])
def test_assign_fp32_parity(dev, backend, metric, N, K, D):
    x, E = O.synthetic_latents(N, K, D, seed=3407 + N + D)
    q_ref, d = O.encode(metric, x, E)
    q, score = _assign(x, E, metric, dev, backend)
    n_diff = _check_indices(d, q_ref, q.cpu(), what=f'{metric} {N}x{K}x{D}')
    assert n_diff <= max(2, N // 200)
    assert int(q.min()) >= 0 and int(q.max()) < K
    # the kernel's score is <x, e> - 0.5|e|^2 (L2) or <x, e/|e|> (cosine)
    Ef = F.normalize(E) if metric == 'Cosine' else E
    s_ref = (x * Ef[q.cpu()]).sum(1) - (0.5 * (Ef[q.cpu()] ** 2).sum(1) if metric == 'L2' else 0)
    torch.testing.assert_close(score.cpu(), s_ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('backend', [ops.BACKEND_SIMT, ops.BACKEND_TCGEN05], ids=['simt', 'tcgen05'])
def test_assign_nan_token_rows_return_index_zero(dev, backend):
    """A token with a NaN component has NaN distance to every code: torch.argmin returns 0 for such a row
    (what the reference's `distance.argmin(-1)` yields); the other rows are unaffected and nothing reads out
    of bounds downstream."""
    N, K, D = 300, 77, 32
    x, E = O.synthetic_latents(N, K, D, seed=5)
    x[17, 3] = float('nan')
    x[255] = float('nan')
    q_ref, d = O.encode('L2', x, E)
    assert q_ref[17] == 0 and q_ref[255] == 0
    q, _ = _assign(x, E, 'L2', dev, backend)
    assert q[17] == 0 and q[255] == 0
    keep = torch.ones(N, dtype=torch.bool)
    keep[[17, 255]] = False
    _check_indices(d[keep], q_ref[keep], q.cpu()[keep], what='rows next to NaN rows')


@pytest.mark.parametrize('backend', [ops.BACKEND_SIMT, ops.BACKEND_TCGEN05], ids=['simt', 'tcgen05'])
def test_assign_bf16_tokens_exact_operands(dev, backend):
    """bf16 tokens are ONE exact plane; the fp32 codebook is three: products are exact, so the only
    difference from the fp32 oracle on the up-cast inputs is fp32 accumulation order."""
    N, K, D = 4096, 2048, 32
    x, E = O.synthetic_latents(N, K, D, normalized_codebook=True)
    xb = x.to(torch.bfloat16)
    q_ref, d = O.encode('Cosine', xb.float(), E)
    q, _ = _assign(xb, E, 'Cosine', dev, backend)
    _check_indices(d, q_ref, q.cpu(), what='bf16 tokens')


def test_pack_rows_fp16_pair_format(dev):
    """VQB_PLANES_F16X2: hi = fp16(v), lo' = fp16((v - hi) * 2^11) of the NORMALISED row;
    hi + lo' * 2^-11 reproduces v to 22 bits (abs error <= max(2^-22 |v|, 2^-36))."""
    g = torch.Generator().manual_seed(7)
    for rows, D, dtype in ((1000, 32, torch.float32), (300, 256, torch.float32), (77, 20, torch.float32),
                           (513, 768, torch.bfloat16)):
        src = (torch.randn(rows, D, generator=g) * 3).to(dtype)
        wb = torch.empty(rows, D, dtype=torch.float32, device=dev)
        op = ops.pack_rows(src.to(dev), normalize=True, fmt='f16x2', writeback=wb)
        assert op.pair and op.nplanes == 2 and op.abi_planes == 0x12
        planes = op.planes.view(torch.float16).float().cpu()
        v = wb.cpu()                                # the kernel's own fp32 F.normalize(row)
        torch.testing.assert_close(v, F.normalize(src.float()), rtol=3e-7, atol=1e-9)
        rec = (planes[0, :rows, :D].double() + planes[1, :rows, :D].double() / 2048)
        err = (rec - v.double()).abs()
        assert (err <= (v.abs().double() * 2.0 ** -22).clamp_min(2.0 ** -36) + 1e-12).all(), float(err.max())
        assert not planes[:, rows:].any() and not planes[:, :, D:].any()                    # padding zeroed
    with pytest.raises(ValueError):
        ops.pack_rows(src.to(dev), normalize=False, fmt='f16x2')


@pytest.mark.parametrize('backend', [ops.BACKEND_SIMT, ops.BACKEND_TCGEN05], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('N,K,D', [(4096, 2048, 32), (1000, 700, 64), (384, 1100, 256), (257, 4100, 20), (300, 333, 768),
                                   (2048, 512, 8)])
def test_assign_fp16_pair_codebook(dev, backend, N, K, D):
    """bf16 tokens (one plane) x fp16-pair codebook: two MMA terms, the lo' term first and the accumulator scaled
    by 2^-11 (scale-input-d) when the hi term is added.  Indices equal the fp32 oracle except near-ties; the
    score is <x, e/|e|> to fp32 rounding."""
    x, E = O.synthetic_latents(N, K, D, seed=11 + D)
    xb = x.to(torch.bfloat16)
    q_ref, d = O.encode('Cosine', xb.float(), E)
    book = ops.pack_rows(E.to(dev), normalize=True, fmt='f16x2')
    toks = ops.pack_rows(xb.to(dev), fmt='f16')
    keys = ops.new_keys(N, dev)
    ops.assign(toks, book, keys, l2=False, backend=backend)
    q, score = ops.unpack_keys(keys, want_score=True)
    _check_indices(d, q_ref, q.cpu(), what=f'fp16 pair {N}x{K}x{D}')
    s_ref = (xb.double() * F.normalize(E).double()[q.cpu()]).sum(1)
    scale = xb.float().norm(dim=1).double()
    # 2^-22 operand representation + fp32 accumulation over D products
    tol = (2.0 ** -22 + 2.0 ** -23 * D ** 0.5) * scale + 1e-9
    assert ((score.cpu().double() - s_ref).abs() <= tol).all()


def test_assign_fp16_pair_both_operands_column_argmin(dev):
    """pair x pair (three terms: a_lo'.b_hi, a_hi.b_lo', then the scaled a_hi.b_hi): the column arg-min of
    NearestAnchor with normalised tokens that are not a zero-copy operand (D = 20 needs padding)."""
    N, K, D = 1500, 300, 20
    x, E = O.synthetic_latents(N, K, D, seed=5)
    xb = x.to(torch.bfloat16)
    d = 1 - F.normalize(xb.float()) @ F.normalize(E).t()
    book = ops.pack_rows(E.to(dev), normalize=True, fmt='f16x2')
    toks = ops.pack_rows(xb.to(dev), normalize=True, fmt='f16x2')
    for backend in (ops.BACKEND_SIMT, ops.BACKEND_TCGEN05):
        keys = ops.new_keys(K, dev)
        ops.assign(book, toks, keys, l2=False, backend=backend)
        _check_indices(d.t().contiguous(), d.argmin(0), ops.unpack_keys(keys).cpu(), what='pair x pair')


def test_column_argmin_fp16_token_plane_with_scaled_rows(dev):
    """NearestAnchor column arg-min against a fp16-pair codebook with bf16 tokens: tokens as ONE fp16 plane +
    1/|x_n| column scale (two MMA terms).  A token with huge components is stored scaled by a power of two; the
    same factor is folded into its inverse norm, so it competes with its true cosine."""
    from vector_quantization_b200 import functional as Fq
    N, K, D = 3000, 257, 32
    x, E = O.synthetic_latents(N, K, D, seed=9)
    x[5] = E[100] * 1e6                 # exactly aligned with code 100, far outside the fp16 range
    x[7] = E[200] * 3e-4
    xb = x.to(torch.bfloat16)
    d = 1 - F.normalize(xb.float()) @ F.normalize(E).t()
    book = Fq.pack_codebook(E.to(dev), 'Cosine', tokens=xb.to(dev))
    assert book.pair
    for backend in (ops.BACKEND_SIMT, ops.BACKEND_TCGEN05):
        raw = ops.pack_rows(xb.to(dev), fmt='f16')
        raw.inv_norm = ops.row_inv_norm(xb.to(dev), f16_rows=True)
        keys = ops.new_keys(K, dev)
        ops.assign(book, raw, keys, l2=False, scale_columns=True, backend=backend)
        idx, score = ops.unpack_keys(keys, want_score=True)
        _check_indices(d.t().contiguous(), d.argmin(0), idx.cpu(), what='fp16 token plane column arg-min')
        assert idx[100] == 5 and idx[200] == 7
        torch.testing.assert_close(score.cpu(), (1 - d).max(0).values, rtol=1e-5, atol=2e-6)
    keys2 = Fq.column_nearest(xb.to(dev), book, 'Cosine')
    assert torch.equal(ops.unpack_keys(keys2), idx)


def test_assign_rejects_mixed_fp16_bf16_operands(dev):
    """kind::f16 MMAs take fp16 x fp16 or bf16 x bf16 (a mixed pair is an illegal instruction on B200).  The one
    supported mix is a ONE-plane bf16 A operand with D <= 64, converted in shared memory by the kernel."""
    from vector_quantization_b200._lib import VQBError
    x, E = O.synthetic_latents(300, 64, 32)
    book = ops.pack_rows(E.to(dev), normalize=True, fmt='f16x2')
    with pytest.raises(VQBError):
        ops.assign(ops.pack_rows(x.to(dev), planes=3), book, ops.new_keys(300, dev), l2=False)
    x2, E2 = O.synthetic_latents(300, 64, 128)
    with pytest.raises(VQBError):   # D > 64: k-blocked ring, no resident token tile to convert
        ops.assign(ops.pack_rows(x2.to(torch.bfloat16).to(dev), planes=1),
                   ops.pack_rows(E2.to(dev), normalize=True, fmt='f16x2'), ops.new_keys(300, dev), l2=False)


@pytest.mark.parametrize('backend', [ops.BACKEND_SIMT, ops.BACKEND_TCGEN05], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('N,K,D', [(4096, 2048, 32), (1000, 700, 64), (333, 5000, 16), (65536, 1024, 32)])
def test_assign_zero_copy_bf16_tokens_against_fp16_pair(dev, backend, N, K, D):
    """Raw bf16 tokens handed over zero-copy against a fp16-pair codebook: the kernel converts the resident token
    tile to fp16 in shared memory.  Same keys as with a packed fp16 token plane, and parity with the oracle."""
    x, E = O.synthetic_latents(N, K, D, seed=3 + D)
    xb = x.to(torch.bfloat16).to(dev)
    book = ops.pack_rows(E.to(dev), normalize=True, fmt='f16x2')
    raw = ops.as_operand(xb)
    assert raw is not None and raw.fmt == 'bf16' and raw.planes.data_ptr() == xb.data_ptr()
    k_raw, k_packed = ops.new_keys(N, dev), ops.new_keys(N, dev)
    ops.assign(raw, book, k_raw, l2=False, backend=backend)
    ops.assign(ops.pack_rows(xb, fmt='f16'), book, k_packed, l2=False, backend=backend)
    if backend == ops.BACKEND_TCGEN05:
        assert torch.equal(k_raw, k_packed)
    q_ref, d = O.encode('Cosine', xb.float().cpu(), E)
    _check_indices(d, q_ref, ops.unpack_keys(k_raw).cpu(), what=f'zero-copy bf16 x pair {N}x{K}x{D}')


def test_pack_rows_fp16_single_plane(dev):
    """VQB_PLANES_F16: bf16 tokens are exact in one fp16 plane; a row with a huge component is scaled by a power
    of two (its arg-max is unchanged)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(515, 24, generator=g).to(torch.bfloat16)
    x[7] *= 1e6                                   # beyond the fp16 range
    x[9, 0] = 3e-6                                # below the exact range: abs error <= 2^-25
    op = ops.pack_rows(x.to(dev), fmt='f16')
    assert op.fmt == 'f16' and op.abi_planes == 0x11 and op.nplanes == 1
    got = op.planes.view(torch.float16)[0, :515, :24].float().cpu()
    keep = torch.ones(515, dtype=torch.bool); keep[[7, 9]] = False
    assert torch.equal(got[keep], x.float()[keep])
    ratio = (got[7] / x.float()[7])
    assert torch.isfinite(got[7]).all() and got[7].abs().max() < 2 ** 15
    assert (ratio == ratio[0]).all() and float(torch.log2(ratio[0])) == round(float(torch.log2(ratio[0])))
    assert (got[9] - x.float()[9]).abs().max() <= 2.0 ** -25
    with pytest.raises(ValueError):
        ops.pack_rows(x.float().to(dev), fmt='f16')


def _random_shapes(n, seed):
    g = torch.Generator().manual_seed(seed)
    dims = [1, 2, 3, 5, 8, 12, 16, 20, 31, 32, 33, 48, 64, 65, 96, 128, 200, 256, 320]
    out = []
    for _ in range(n):
        N = int(torch.randint(1, 1500, (1,), generator=g))
        K = int(torch.randint(1, 3000, (1,), generator=g))
        D = dims[int(torch.randint(0, len(dims), (1,), generator=g))]
        out.append((N, K, D))
    return out


@pytest.mark.parametrize('N,K,D', _random_shapes(24, seed=2024))
def test_assign_random_shapes_all_formats(dev, N, K, D):
    """Seeded random (ragged) shapes: every operand format of the tcgen05 kernel against the fp32 oracle —
    fp32 tokens as three bf16 planes (L2 and cosine), bf16 tokens against the fp16-pair codebook (packed fp16
    plane, and zero-copy with in-kernel conversion where the layout allows it)."""
    x, E = O.synthetic_latents(N, K, D, seed=N * 7 + K)
    for metric in ('L2', 'Cosine'):
        q_ref, d = O.encode(metric, x, E)
        q, _ = _assign(x, E, metric, dev, ops.BACKEND_TCGEN05)
        _check_indices(d, q_ref, q.cpu(), what=f'{metric} bf16 planes {N}x{K}x{D}', squared=metric == 'L2')
    xb = x.to(torch.bfloat16)
    q_ref, d = O.encode('Cosine', xb.float(), E)
    book = ops.pack_rows(E.to(dev), normalize=True, fmt='f16x2')
    variants = [ops.pack_rows(xb.to(dev), fmt='f16')]
    raw = ops.as_operand(xb.to(dev))
    if raw is not None and D <= 64:
        variants.append(raw)
    for toks in variants:
        keys = ops.new_keys(N, dev)
        ops.assign(toks, book, keys, l2=False)
        _check_indices(d, q_ref, ops.unpack_keys(keys).cpu(), what=f'pair codebook ({toks.fmt} tokens) {N}x{K}x{D}')


def test_assign_tie_break_lowest_index(dev):
    """Exact ties resolve to the lowest index, like torch.argmin (duplicate codebook rows)."""
    N, D = 512, 32
    g = torch.Generator().manual_seed(5)
    base = torch.randn(300, D, generator=g).to(torch.bfloat16).float()
    E = torch.cat([base, base, base])            # every code appears three times: 300-periodic ties
    x = base[torch.randint(0, 300, (N,), generator=g)] + 0.01 * torch.randn(N, D, generator=g)
    x = x.to(torch.bfloat16)
    for backend in (ops.BACKEND_SIMT, ops.BACKEND_TCGEN05):
        q, _ = _assign(x, E, 'L2', dev, backend)
        assert int(q.max()) < 300, 'a duplicate with a higher index won a tie'
        q_ref, d = O.encode('L2', x.float(), E)
        _check_indices(d, q_ref, q.cpu(), what='ties')


def test_assign_column_argmin_swapped_operands(dev):
    """NearestAnchor's d.argmin(0) == the same kernel with codes as rows and tokens as columns."""
    N, K, D = 3000, 640, 32
    x, E = O.synthetic_latents(N, K, D, seed=11)
    for metric in ('L2', 'Cosine'):
        _, d = O.encode(metric, x, E)
        idx_ref = d.argmin(0)
        cos = metric == 'Cosine'
        codes = ops.pack_rows(E.to(dev), normalize=cos)
        toks = ops.pack_rows(x.to(dev), normalize=cos, want_half_sqnorm=not cos)
        keys = ops.new_keys(K, dev)
        ops.assign(codes, toks, keys, l2=not cos)
        idx = ops.unpack_keys(keys).cpu()
        _check_indices(d.t().contiguous(), idx_ref, idx, what=f'column argmin {metric}')


def test_assign_column_argmin_raw_tokens_with_column_scale(dev):
    """Cosine column arg-min with RAW bf16 tokens (one exact plane, zero-copy) and the 1/|x_n| column scale applied
    in the epilogue == arg-min over the normalised tokens (what NearestAnchor needs)."""
    from vector_quantization_b200 import functional as Fq
    N, K, D = 5000, 700, 32
    x, E = O.synthetic_latents(N, K, D, seed=13)
    x = (x * (0.5 + torch.rand(N, 1, generator=torch.Generator().manual_seed(1)) * 4)).to(torch.bfloat16)  # varied norms
    _, d = O.encode('Cosine', x.float(), E)
    idx_ref = d.argmin(0)
    book = ops.pack_rows(E.to(dev), normalize=True)
    for backend in (ops.BACKEND_TCGEN05, ops.BACKEND_SIMT):
        raw = ops.as_operand(x.to(dev))
        raw.inv_norm = ops.row_inv_norm(x.to(dev))
        keys = ops.new_keys(K, dev)
        ops.assign(book, raw, keys, l2=False, scale_columns=True, backend=backend)
        _check_indices(d.t().contiguous(), idx_ref, ops.unpack_keys(keys).cpu(), what='raw-token column argmin')
    keys = Fq.column_nearest(x.to(dev), book, 'Cosine')
    _check_indices(d.t().contiguous(), idx_ref, ops.unpack_keys(keys).cpu(), what='column_nearest')
    inv = ops.row_inv_norm(x.to(dev)).cpu()
    torch.testing.assert_close(inv[:N], 1 / x.float().norm(dim=1), rtol=2e-6, atol=0)
    assert (inv[N:] == 0).all()


def test_assign_sharded_codebook_min_combine(dev):
    """Two launches over two codebook shards min-combine into the same keys as one launch."""
    N, K, D = 2000, 1024, 32
    x, E = O.synthetic_latents(N, K, D, seed=21)
    q_ref, d = O.encode('L2', x, E)
    a = ops.pack_rows(x.to(dev))
    keys = ops.new_keys(N, dev)
    for r in range(2):
        b = ops.pack_rows(E[r * 512:(r + 1) * 512].to(dev), want_half_sqnorm=True)
        ops.assign(a, b, keys, l2=True, index_offset=r * 512)
    _check_indices(d, q_ref, ops.unpack_keys(keys).cpu(), what='sharded')


# --------------------------------------------------------------------------------------------


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('normalize_x', [False, True], ids=['raw', 'normx'])
@pytest.mark.parametrize('N,K,D,norm', [(1000, 64, 32, False), (777, 100, 256, True), (4096, 512, 8, True),
                                        (333, 50, 20, False), (130, 40, 768, True), (257, 30, 5, True)])
def test_gather_ste_loss_and_backward(dev, dtype, normalize_x, N, K, D, norm):
    """Fused forward (token normalise + key unpack + gather + STE + MSE terms) and closed-form backward
    (incl. the chain through F.normalize) against autograd on the oracle."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, D, generator=g).to(dtype)
    W = torch.randn(K, D, generator=g)
    q = torch.randint(0, K, (N,), generator=g)
    gz = torch.randn(N, D, generator=g)
    g4 = torch.tensor([0.7, -1.3, 0.4, 2.0])
    if not norm:
        g4[2:] = 0
    # oracle (fp32 on up-cast inputs), autograd for the gradients
    xin = x.float().clone().requires_grad_(True)
    xo = O.normalize(xin) if normalize_x else xin
    Wo = W.clone().requires_grad_(True)
    z = O.decode(Wo, q)
    cb, cm = O.codebook_loss(z, xo), O.commitment_loss(z, xo)
    cbn, cmn = O.codebook_loss(z, xo, True), O.commitment_loss(z, xo, True)
    z_ste = O.ste(z, xo)
    total = (z_ste * gz).sum() + g4[0] * cb + g4[1] * cm + g4[2] * cbn + g4[3] * cmn
    total.backward()

    xd, Wd, qd = x.to(dev), W.to(dev), q.to(dev)
    from vector_quantization_b200 import parallel
    keys = parallel.pack_keys_host(torch.randn(N, generator=g), q + 7).to(dev)   # packed keys, index offset 7
    z_k, mse4, q_out, xn = ops.gather_ste_loss(xd, Wd, keys=keys, key_offset=7, normalize_x=normalize_x,
                                               want_norm=norm, want_quant=True, want_xnorm=normalize_x)
    assert torch.equal(q_out.cpu(), q)
    if normalize_x:
        torch.testing.assert_close(xn.cpu(), xo.detach(), rtol=2e-6, atol=1e-7)
        torch.testing.assert_close(z_k.cpu(), z_ste.detach(), rtol=1e-5, atol=1e-6)
    else:
        assert torch.equal(z_k.cpu(), z_ste.detach()), 'z_ste must be bit-exact: x + (W[q] - x)'
    ref4 = torch.stack([cb, cm, cbn, cmn]).detach()
    if not norm:
        ref4[2:] = mse4.cpu()[2:]
    torch.testing.assert_close(mse4.cpu(), ref4, rtol=1e-5, atol=1e-8)
    # deterministic reduction: a second launch (indices instead of keys) returns the same bits
    _, mse4b, _, _ = ops.gather_ste_loss(xd, Wd, quant=qd, normalize_x=normalize_x, want_norm=norm)
    assert torch.equal(mse4, mse4b)

    gx, gW = ops.quantize_backward(gz.to(dev), xd, Wd, qd, g4.to(dev), normalize_x=normalize_x, want_norm=norm,
                                   need_gW=True)
    tol = dict(rtol=1e-4, atol=2e-6) if dtype == torch.float32 else dict(rtol=2 ** -7, atol=2e-3)
    torch.testing.assert_close(gx.float().cpu(), xin.grad, **tol)
    torch.testing.assert_close(gW.cpu(), Wo.grad, rtol=1e-4, atol=1e-6)


def test_assign_zero_copy_bf16_tokens(dev):
    """A contiguous bf16 [N, D] tensor with D in {16, 32, 64k} is passed to the TMA descriptor as is."""
    for N, K, D in ((1000, 512, 32), (300, 700, 64), (513, 256, 16)):
        x, E = O.synthetic_latents(N, K, D, seed=N, normalized_codebook=True)
        xb = x.to(torch.bfloat16)
        q_ref, d = O.encode('Cosine', xb.float(), E)
        tok = ops.as_operand(xb.to(dev))
        assert tok is not None and tok.planes.data_ptr() != 0 and tok.plane_rows == N
        book = ops.pack_rows(E.to(dev), normalize=True)
        for backend in (ops.BACKEND_TCGEN05, ops.BACKEND_SIMT):
            keys = ops.new_keys(N, dev)
            ops.assign(tok, book, keys, l2=False, backend=backend)
            _check_indices(d, q_ref, ops.unpack_keys(keys).cpu(), what=f'zero-copy {N}x{K}x{D}')
    assert ops.as_operand(torch.zeros(8, 20, dtype=torch.bfloat16, device=dev)) is None


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('N,D', [(1000, 32), (77, 8), (300, 256), (65, 768), (10, 5)])
def test_l2norm(dev, dtype, N, D):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, D, generator=g).to(dtype)
    x[0] = 0  # eps-clamped row
    xo = x.float().clone().requires_grad_(True)
    y = F.normalize(xo)
    gy = torch.randn(N, D, generator=g)
    y.backward(gy)
    yk = ops.l2norm_forward(x.to(dev))
    torch.testing.assert_close(yk.cpu(), y.detach(), rtol=2e-6, atol=1e-7)
    gx = ops.l2norm_backward(gy.to(dev), x.to(dev))
    tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=2 ** -7, atol=1e-2)
    torch.testing.assert_close(gx.float().cpu()[1:], xo.grad[1:], **tol)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('N,K,D,normalize,hot', [(5000, 300, 32, True, False), (4096, 64, 8, False, True),
                                                 (1000, 128, 256, True, False), (999, 50, 20, False, True),
                                                 (3000, 16, 768, True, True)])
def test_scatter_stats_and_kmeans_update(dev, dtype, N, K, D, normalize, hot):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, D, generator=g).to(dtype)
    q = torch.randint(0, K // 2 if hot else K, (N,), generator=g)  # hot: half the codebook unused
    if hot:
        q[: N // 2] = 3  # long runs of the same code exercise the warp-aggregated path
    W = F.normalize(torch.randn(K, D, generator=g))
    xs = F.normalize(x.float()) if normalize else x.float()
    sums = torch.zeros(K, D).index_add_(0, q, xs)
    cnt = q.bincount(minlength=K)
    stats = ops.scatter_stats(x.to(dev), q.to(dev), K, normalize_x=normalize)
    assert torch.equal(stats[K * D:].cpu(), cnt.float()), 'counts are exact'
    torch.testing.assert_close(stats[:K * D].view(K, D).cpu(), sums, rtol=1e-4, atol=2e-5)
    counts = torch.zeros(K, dtype=torch.int64, device=dev)
    ops.bincount_accumulate(q.to(dev), counts)
    ops.bincount_accumulate(q.to(dev), counts)
    assert torch.equal(counts.cpu(), 2 * cnt)
    if normalize:  # the VQ-KD update on top of these statistics
        W_ref = O.vqkd_update([x.float()], [q], W, 0.99)
        Wd = W.to(dev).clone()
        ops.kmeans_ema_update(stats, Wd, 0.99)
        torch.testing.assert_close(Wd.cpu(), W_ref, rtol=1e-5, atol=1e-6)


def test_cvq_update(dev):
    N, K, D = 3000, 256, 32
    x, E = O.synthetic_latents(N, K, D, seed=5)
    q, d = O.encode('Cosine', x, E)
    prob = torch.rand(K, generator=torch.Generator().manual_seed(2)) / K
    W_ref, p_ref, anchors_ref, idx_ref = O.cvq_update([x], [d], [q], E, prob, 0.99, 1e-3, False)
    idx_ref = idx_ref[0]
    xd = x.to(dev)
    codes = ops.pack_rows(E.to(dev), normalize=True)
    toks = ops.pack_rows(xd, normalize=True)
    keys = ops.new_keys(K, dev)
    ops.assign(codes, toks, keys, l2=False)
    _check_indices(d.t().contiguous(), idx_ref, ops.unpack_keys(keys).cpu(), what='anchor index')
    anchors = ops.gather_rows_by_key(xd, keys)
    same = ops.unpack_keys(keys).cpu() == idx_ref
    assert torch.equal(anchors.cpu()[same], anchors_ref[same])
    Wd, pd = E.to(dev).clone(), prob.to(dev).clone()
    counts = q.bincount(minlength=K).to(dev)
    ops.cvq_update(Wd, anchors_ref.to(dev), pd, counts, torch.tensor([N], device=dev), decay=0.99, eps=1e-3)
    assert torch.equal(ops.embedding_gather(E.to(dev), q.to(dev).view(30, 100)).cpu(), E[q].view(30, 100, D))
    torch.testing.assert_close(pd.cpu(), p_ref, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(Wd.cpu(), W_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('levels', [[8, 8, 5, 5, 5], [8, 8, 8, 5, 5, 5], [7, 5, 5, 5, 5], [4, 3]])
@pytest.mark.parametrize('N', [1, 255, 24576])
def test_fsq(dev, levels, N):
    fsq = O.FSQ(levels)
    g = torch.Generator().manual_seed(N)
    x = 1.5 * torch.randn(N, len(levels), generator=g)
    xo = x.clone().requires_grad_(True)
    quant, zq, pre = fsq.encode(xo)
    gz = torch.randn(N, len(levels), generator=g)
    zq.backward(gz)
    p = ops.fsq_params(levels, 1e-3)
    zk, ik = ops.fsq_forward(x.to(dev), p)
    # round-half-even boundary: tanhf differs from the CPU tanh by an ulp or two
    safe = ((pre.detach() - pre.detach().floor() - 0.5).abs() > 1e-5).all(1)
    assert safe.float().mean() > 0.99
    assert torch.equal(ik.cpu()[safe], quant[safe])
    assert torch.equal(zk.cpu()[safe], zq.detach()[safe])
    assert ik.dtype == torch.int32 and int(ik.max()) < fsq.codebook_size and int(ik.min()) >= 0
    gx = ops.fsq_backward(gz.to(dev), x.to(dev), p)
    torch.testing.assert_close(gx.cpu(), xo.grad, rtol=1e-4, atol=1e-6)
    # decode-only path (digit / half - 1) == the reference's decode-only branch, bit for bit; it agrees with the
    # encode path's r / half only up to fp32 rounding (e.g. 4/3 - 1 != 1/3), exactly as in the reference
    zd = ops.fsq_decode(ik, p)
    assert torch.equal(zd.cpu(), fsq.decode(ik.cpu().long()).float())
    torch.testing.assert_close(zd, zk, rtol=0, atol=2e-7)


# ---- certified one-term pass (D >= 128, fp16-pair codebook): identical to the two-term contraction -----------------
@pytest.mark.parametrize('clustered', [True, False], ids=['clustered', 'random'])
@pytest.mark.parametrize('N,K,D', [(4096, 2048, 128), (3000, 1000, 256), (2048, 4096, 768), (513, 300, 200)])
def test_certified_one_term_assignment_equals_two_term(dev, N, K, D, clustered):
    """ONE MMA term (hi plane) + certificate + exact re-run of the uncertified rows == the two-term fp16-pair
    contraction, index for index, for the row arg-min (tokens x codebook) and the column arg-min (codebook x tokens
    with the 1/|x| column scale).  On random data a few per cent of the rows fail the certificate and take the
    re-run; on clustered data almost none do."""
    from vector_quantization_b200 import functional as Fq
    x, E = O.synthetic_latents(N, K, D, seed=N + D, normalized_codebook=True, clustered=clustered)
    xb = x.to(torch.bfloat16).to(dev)
    book = ops.pack_rows(E.to(dev), normalize=True, fmt='f16x2', want_lo_norm=True)
    assert 0 < float(book.lo_norm_max) < 2.0 ** -11
    toks = ops.pack_rows(xb, fmt='f16')
    toks.inv_norm = ops.row_inv_norm(xb, f16_rows=True)
    # row arg-min
    want = ops.assign(toks, book, ops.new_keys(N, dev), l2=False)
    got = Fq.certified_assign(toks, book, ops.new_keys(N, dev), a_inv_norm=toks.inv_norm)
    flagged_rows = int(Fq.LAST_CERTIFY['count'])
    assert torch.equal(ops.unpack_keys(got), ops.unpack_keys(want))
    # column arg-min (NearestAnchor): codes as rows, raw tokens as columns with the 1/|x_n| scale
    want_c = ops.assign(book, toks, ops.new_keys(K, dev), l2=False, scale_columns=True)
    got_c = Fq.certified_assign(book, toks, ops.new_keys(K, dev), scale_columns=True)
    flagged_cols = int(Fq.LAST_CERTIFY['count'])
    assert torch.equal(ops.unpack_keys(got_c), ops.unpack_keys(want_c))
    if clustered:
        assert flagged_rows <= N // 50
    else:
        assert 0 < flagged_rows < N // 2 and 0 < flagged_cols < K, (flagged_rows, flagged_cols)
    # ... and it is the oracle's arg-min under the near-tie policy
    q_ref, d = O.encode('Cosine', xb.float().cpu(), E)
    rows, gap = O.index_mismatch_report(d, q_ref, ops.unpack_keys(got).cpu())
    assert (gap < 1e-5).all() and rows.numel() <= max(2, N // 200)


def test_certified_pass_through_the_module(dev):
    """cfg-4-like CVQ-VAE step (D = 256, cosine, bf16 tokens): the module takes the certified path for both passes and
    still matches the oracle step."""
    from vector_quantization_b200 import functional as Fq
    import vector_quantization_b200 as vqb
    N, K, D = 2048, 512, 256
    x, E = O.synthetic_latents(N, K, D, seed=21, normalized_codebook=True)
    cfg = dict(type='VQGANQuantizer', embedding=dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=K, embedding_dim=D),
               distance=dict(type='CosineDistance'),
               callbacks=[dict(type='CVQVAECallback', ema=dict(), anchor=dict(type='NearestAnchor'))],
               losses=dict(vqgan_loss=dict(type='VQGANLoss')), init_weights=dict(type='vqgan'))
    q = vqb.build_quantizer(cfg, training=True).to(dev)
    with torch.no_grad():
        q.embedding.weight.copy_(E)
    Fq.LAST_CERTIFY.clear()
    xb = x.to(torch.bfloat16)
    z, loss, memo = q(xb.to(dev).requires_grad_(True), dict())
    assert 'count' in Fq.LAST_CERTIFY, 'the certified one-term pass did not run'
    spec = O.QuantizerSpec(distance='Cosine', callback='CVQVAECallback', losses={'vqgan_loss': dict(type='VQGANLoss')})
    out = O.quantizer_forward(spec, [xb.float()], E, torch.zeros(K))
    rows, gap = O.index_mismatch_report(out['distance'][0], out['quant'][0], memo['quant'].cpu())
    assert (gap < 1e-5).all() and rows.numel() <= 4
    torch.testing.assert_close(loss.detach().cpu(), out['loss'][0].detach(), rtol=1e-5, atol=1e-7)
    if rows.numel() == 0:
        ck = memo['encode']['column_keys'].cpu()
        has_key = ck != -1                      # the column pass is restricted to the codes whose anchor has weight
        col = ops.unpack_keys(memo['encode']['column_keys']).cpu()
        same = (col == out['anchor_idx'][0]) | ~has_key
        assert has_key.any() and same.float().mean() > 0.99
        torch.testing.assert_close(q.embedding.weight.detach().cpu()[same], out['weight'][same], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(q.get_buffer('_probability').cpu(), out['prob'], rtol=1e-6, atol=1e-9)


def test_cvq_needy_codes_is_a_superset_of_the_codes_with_nonzero_anchor_weight(dev):
    """vqb_cvq_needy_codes: ascending list of the codes whose fp32 blend weight 1 - decay_k can be non-zero, from a
    LOWER bound of the new probability (local counts over the global total).  It must contain every code whose true
    weight (global counts) is non-zero, and leave out the clearly busy ones."""
    K, N, world = 4096, 16384, 4
    g = torch.Generator().manual_seed(5)
    prob = torch.rand(K, generator=g) / K
    prob[::3] *= 1e-4
    local = torch.randint(0, 6, (K,), generator=g)
    others = torch.randint(0, 14, (K,), generator=g)
    cnt_local = torch.cat([local, torch.tensor([N])]).to(torch.int64)
    code_list, count, compact = ops.cvq_needy_codes(prob.to(dev), cnt_local.to(dev), world * N, decay=0.99, eps=1e-3)
    n = int(count)
    listed = code_list[:n].cpu().long()
    assert torch.equal(listed, listed.sort().values) and listed.unique().numel() == n
    assert (compact[:n].cpu() == -1).all()
    p_true = O.ema(prob, (local + others).float() / (world * N), 0.99)
    weight = 1 - (1 - torch.exp(-p_true * K * 10 / (1 - 0.99) - 1e-3))
    needed = (weight != 0).nonzero().flatten()
    mask = torch.zeros(K, dtype=torch.bool)
    mask[listed] = True
    assert mask[needed].all(), 'a code with a non-zero anchor weight is missing from the list'
    assert 0.1 < mask.float().mean() < 0.9


@pytest.mark.parametrize('N,K,D,x_dtype,e_dtype', [
    (3000, 1000, 8, torch.float32, torch.float32),      # 3 x 3 planes: the pieces of -0.5|e|^2 sit in one column, three planes
    (2500, 700, 8, torch.bfloat16, torch.float32),      # 1 x 3
    (2500, 700, 10, torch.float32, torch.bfloat16),     # 3 x 1: the pieces are spread over three columns of plane 0
    (4096, 16384 // 8, 8, torch.bfloat16, torch.bfloat16),
    (1000, 300, 24, torch.float32, torch.float32),      # Dp = 32
])
def test_assign_l2_side_terms_folded_into_the_contraction(dev, N, K, D, x_dtype, e_dtype):
    """vqb_fold_l2_side: with the L2 side terms written into the spare operand columns, the plain arg-max of the
    contraction (side_mode 0) is the arg-min of the reference's cdist (vq/algorithms/vq/distances.py:28-35), for the
    row pass (nearest code per token) and the column pass (nearest token per code), on both backends."""
    from vector_quantization_b200 import functional as Fq
    assert ops.can_fold_l2(D)
    x, E = O.synthetic_latents(N, K, D, seed=21)
    x, E = x.to(x_dtype), E.to(e_dtype)
    q_ref, d = O.encode('L2', x.float(), E.float())
    book = Fq.pack_codebook(E.to(dev).float() if e_dtype == torch.float32 else E.to(dev), 'L2')
    assert book.folded == 'codes'
    keys = Fq.nearest_code(x.to(dev), book, 'L2')
    q, score = ops.unpack_keys(keys, want_score=True)
    _check_indices(d, q_ref, q.cpu(), what='folded L2 row arg-min', squared=True)
    # the folded score is -0.5 * |x - e|^2 up to the 1-column products (tokens role without its own term: + 0.5|x|^2)
    want = -0.5 * (d[torch.arange(N), q.cpu()] ** 2) + 0.5 * x.float().pow(2).sum(1)
    assert torch.allclose(score.cpu(), want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))
    col = ops.unpack_keys(Fq.column_nearest(x.to(dev), book, 'L2')).cpu()
    _check_indices(d.t().contiguous(), d.argmin(0), col, what='folded L2 column arg-min', squared=True)
    # same answers as the un-folded side_mode-1 path and as the SIMT backend on the folded operands
    a = ops.pack_rows(x.to(dev), planes=None)
    b = ops.pack_rows(E.to(dev), want_half_sqnorm=True)
    plain = ops.unpack_keys(ops.assign(a, b, ops.new_keys(N, dev), l2=True)).cpu()
    _check_indices(d, plain, q.cpu(), what='folded vs side-term L2', squared=True)
    ops.fold_l2_side(a, 'tokens')
    ops.fold_l2_side(b, 'codes')
    # the fold fused into the pack launch (vqb_pack_rows_fold) writes the same planes as pack + vqb_fold_l2_side
    assert torch.equal(ops.pack_rows(x.to(dev), planes=None, fold='tokens').planes, a.planes)
    assert torch.equal(ops.pack_rows(E.to(dev), want_half_sqnorm=True, fold='codes').planes, b.planes)
    with_h = ops.fold_l2_side(ops.pack_rows(x.to(dev), planes=None, want_half_sqnorm=True), 'tokens')
    assert torch.equal(ops.pack_rows(x.to(dev), planes=None, want_half_sqnorm=True, fold='tokens+h').planes, with_h.planes)
    simt = ops.unpack_keys(ops.assign(a, b, ops.new_keys(N, dev), l2=True, backend=ops.BACKEND_SIMT)).cpu()
    _check_indices(d, simt, q.cpu(), what='folded L2, SIMT backend', squared=True)


def test_fold_l2_side_rejects_operands_without_spare_columns(dev):
    x = torch.randn(64, 32, device=dev)
    op = ops.pack_rows(x, want_half_sqnorm=True)
    assert not ops.can_fold_l2(32)
    with pytest.raises(Exception):
        ops.fold_l2_side(op, 'codes')
