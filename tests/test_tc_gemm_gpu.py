"""GPU numerics: the tcgen05 dense-layer kernel (csrc/tc_gemm.cu, 3xTF32 split) against a float64
reference of nn.Linear + LeakyReLU (src/model.py:17-30) and of its input gradient.  Tolerance:
max-norm relative error <= 1e-5 (observed 2e-6; the fp32 FFMA tiles give 5e-7)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(engine, mode, x, w, b, act):
    import torch
    from gcrl_b200._lib import check, lib, vp
    M, K = x.shape
    N = w.shape[0]
    y = torch.full((M, N), float("nan"), device="cuda")
    check(lib.gcrl_dense_layer(0, engine, mode, M, N, K, vp(x.data_ptr()), x.stride(0), vp(w.data_ptr()), w.stride(0),
                               None if b is None else vp(b.data_ptr()), None if act is None else vp(act.data_ptr()),
                               0 if act is None else act.stride(0), vp(y.data_ptr()), y.stride(0),
                               vp(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (1, 64, 64), (1000, 256, 256), (65536, 256, 256), (300, 64, 64),
                                   (4096, 512, 512), (777, 96, 100), (5000, 256, 24), (129, 16, 4)])
def test_tc_dense_matches_float64(M, N, K):
    import torch
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    act = torch.randn(M, N, device="cuda")
    ref = x.double() @ w.double().T
    wants = {0: torch.nn.functional.leaky_relu(ref + b.double(), 0.01),
             1: ref * torch.where(act > 0, 1.0, 0.01).double(),
             2: ref + b.double()}
    for mode, want in wants.items():
        got = _run(1, mode, x, w, b if mode != 1 else None, act if mode == 1 else None)
        assert torch.isfinite(got).all()
        err = ((got.double() - want).abs().max() / want.abs().max()).item()
        assert err <= 1e-5, (mode, err)


def test_tc_dense_strided_operands_and_zero_padding():
    """Leading dimensions larger than the logical widths (the agent's padded rows)."""
    import torch
    torch.manual_seed(3)
    M, N, K = 2048, 64, 24
    xs = torch.randn(M, 28, device="cuda")
    ws = torch.randn(N, 28, device="cuda")
    b = torch.randn(N, device="cuda")
    ys = torch.zeros(M, 80, device="cuda")
    from gcrl_b200._lib import check, lib, vp
    check(lib.gcrl_dense_layer(0, 1, 2, M, N, K, vp(xs.data_ptr()), 28, vp(ws.data_ptr()), 28, vp(b.data_ptr()), None, 0,
                               vp(ys.data_ptr()), 80, vp(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    want = xs[:, :K].double() @ ws[:, :K].double().T + b.double()
    assert ((ys[:, :N].double() - want).abs().max() / want.abs().max()).item() <= 1e-5
    assert float(ys[:, N:].abs().max()) == 0.0          # columns beyond N untouched


@pytest.mark.parametrize("M,N,K", [(1024, 128, 32), (4096, 256, 256), (5000, 64, 64), (3000, 256, 24), (8192, 512, 512),
                                   (777, 96, 100), (2048, 4, 64), (16, 256, 256), (65536, 256, 256)])
def test_tc_wgrad_matches_float64(M, N, K):
    """Tensor-core weight gradient (MN-major operands, 32-byte-atom swizzle, split batch): the summed
    partial slabs against dZ^T X in float64; row padding of every slab exactly zero.  Tolerance 1e-5
    max-norm relative (observed 1e-7 .. 7e-6, growing with the rows accumulated per TMEM tile)."""
    import ctypes as C
    import torch
    from gcrl_b200._lib import check, lib, vp
    torch.manual_seed(M + N + K)
    dz = torch.randn(M, N, device="cuda") / M ** 0.5
    x = torch.randn(M, K, device="cuda")
    ldw = (K + 3) // 4 * 4 + 4
    stride = N * ldw + 8
    pw = torch.full((128, stride), float("nan"), device="cuda")
    pb = torch.full((128, N + 4), float("nan"), device="cuda")
    sp = C.c_int()
    check(lib.gcrl_dense_wgrad(0, 1, M, N, K, vp(dz.data_ptr()), N, vp(x.data_ptr()), K, vp(pw.data_ptr()), ldw, stride,
                               vp(pb.data_ptr()), N + 4, 128, C.byref(sp), vp(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    S = sp.value
    assert 1 <= S <= 128
    W = pw[:S, :N * ldw].double().sum(0).reshape(N, ldw)
    want = dz.double().T @ x.double()
    assert ((W[:, :K] - want).abs().max() / want.abs().max()).item() <= 1e-5
    assert float(W[:, K:].abs().max()) == 0.0
    wb = dz.double().sum(0)
    assert ((pb[:S, :N].double().sum(0) - wb).abs().max() / wb.abs().max()).item() <= 1e-5


@pytest.mark.parametrize("M,N,K", [(65536, 256, 256), (16384, 256, 256), (20000, 256, 256), (16385, 256, 256),
                                   (30000, 512, 512), (14500, 256, 24), (1000, 256, 256), (4096, 512, 512),
                                   (8192, 256, 256), (8000, 256, 32)])
def test_tc_dense_presplit_weights_match_float64(M, N, K):
    """The agent's large-batch path: weights pre-split into TF32 hi / lo halves (gcrl_split_tf32), activations split
    in the kernel.  Large M with N % 256 == 0 runs on CTA pairs (cta_group::2: 256 x 256 tiles, every SM stages half
    of the weight tile; a partly filled last round as two 256 x 128 halves per tile),
    the rest on the single-CTA kernel; same tolerance as the un-split entry point (1e-5 max-norm
    relative), ragged row tails and both N tiles covered."""
    import torch
    from gcrl_b200._lib import check, lib, vp
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    act = torch.randn(M, N, device="cuda")
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    st = vp(torch.cuda.current_stream().cuda_stream)
    check(lib.gcrl_split_tf32(0, vp(w.data_ptr()), vp(hi.data_ptr()), vp(lo.data_ptr()), w.numel(), st))
    assert torch.equal((hi.double() + lo.double()).float(), w) or float((hi + lo - w).abs().max()) <= 2e-7 * float(w.abs().max())
    ref = x.double() @ w.double().T
    wants = {0: torch.nn.functional.leaky_relu(ref + b.double(), 0.01),
             1: ref * torch.where(act > 0, 1.0, 0.01).double(),
             2: ref + b.double()}
    for mode, want in wants.items():
        y = torch.full((M, N), float("nan"), device="cuda")
        check(lib.gcrl_dense_layer_presplit(0, mode, M, N, K, vp(x.data_ptr()), K, vp(hi.data_ptr()), vp(lo.data_ptr()), K,
                                            vp(b.data_ptr()) if mode != 1 else None,
                                            vp(act.data_ptr()) if mode == 1 else None, N if mode == 1 else 0,
                                            vp(y.data_ptr()), N, st))
        torch.cuda.synchronize()
        assert torch.isfinite(y).all(), mode
        err = ((y.double() - want).abs().max() / want.abs().max()).item()
        assert err <= 1e-5, (mode, err)
