"""tcgen05/TMEM/TMA GEMM (rcb_gemm_tc) against an fp64 matmul of the same fp32 inputs.
TF32 operands keep a 10-bit mantissa, so the stated tolerance is 2e-3 of the row/column
norm product (|a|.|b|), versus 1e-4 for the fp32 SIMT engine."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 128), (256, 1056, 1056), (70, 99, 99), (640, 4096, 512),
                                   (5, 33, 33), (1000, 64, 4096)])
def test_gemm_tc_matches_fp64(M, N, K):
    from recombiner_b200 import _lib
    from recombiner_b200._lib import check, ptr, stream
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + 3 * N + K)
    lda, ldb, ldc = (K + 3) // 4 * 4, (K + 3) // 4 * 4, (N + 3) // 4 * 4
    A = torch.zeros(M, lda); A[:, :K] = torch.randn(M, K, generator=g)
    Bt = torch.zeros(N, ldb); Bt[:, :K] = torch.randn(N, K, generator=g)
    bias = torch.randn(16, generator=g)
    Ad, Bd, bd = A.cuda(), Bt.cuda(), bias.cuda()
    Cd = torch.full((M, ldc), 7.0, device="cuda")
    check(lib.rcb_gemm_tc(ptr(Ad), lda, ptr(Bd), ldb, ptr(Cd), ldc, M, N, K, None, 1, 0, 0, stream()))
    torch.cuda.synchronize()
    ref = A[:, :K].double() @ Bt[:, :K].double().t()
    scale = (A[:, :K].double().norm(dim=1)[:, None] * Bt[:, :K].double().norm(dim=1)[None, :]).numpy()
    got = Cd[:, :N].cpu().double().numpy()
    assert np.abs(got - ref.numpy()).max() <= 2e-3 * scale.max()
    assert (np.abs(got - ref.numpy()) / scale).max() < 2e-3
    if ldc > N:
        assert float(Cd[:, N:].min()) == 7.0 and float(Cd[:, N:].max()) == 7.0      # padding untouched
    # epilogue: bias (periodic), leaky-relu, accumulate
    check(lib.rcb_gemm_tc(ptr(Ad), lda, ptr(Bd), ldb, ptr(Cd), ldc, M, N, K, ptr(bd), 16, 1, 0, stream()))
    ref2 = torch.nn.functional.leaky_relu(ref + bias[torch.arange(N) % 16].double(), 0.01).numpy()
    got2 = Cd[:, :N].cpu().double().numpy()
    assert (np.abs(got2 - ref2) / scale).max() < 2e-3
    Cd.fill_(1.0)
    check(lib.rcb_gemm_tc(ptr(Ad), lda, ptr(Bd), ldb, ptr(Cd), ldc, M, N, K, None, 1, 0, 1, stream()))
    got3 = Cd[:, :N].cpu().double().numpy()
    assert (np.abs(got3 - (ref.numpy() + 1.0)) / scale).max() < 2e-3


def test_batched_launch_equals_separate_launches():
    """rcb_gemm_tc_batch (one launch, blockIdx.z = problem) against one rcb_gemm_tc / rcb_gemm_tc_h per problem:
    same kernel body, so the results must be bit-identical -- fp32 (TF32) and fp16 operands, ragged N and K."""
    import ctypes as C
    from recombiner_b200 import _lib
    from recombiner_b200._lib import check, ptr, stream
    from recombiner_b200.engine import FitEngine
    lib = _lib.load()
    M, ld = 300, 3272
    Ns, Ks, offs = [1056, 1056, 1056, 99], [1056, 1056, 1056, 104], [0, 1056, 2112, 3168]
    gen = torch.Generator().manual_seed(11)
    for half in (0, 1):
        dt = torch.float16 if half else torch.float32
        es = 2 if half else 4
        A = torch.randn(M, ld, generator=gen).cuda().to(dt)
        A[:, 3267:] = 0
        Bts = [(torch.randn(n, k, generator=gen) / np.sqrt(k)).cuda().to(dt) for n, k in zip(Ns, Ks)]
        c_one = torch.zeros(M, 3268, device="cuda")
        c_bat = torch.zeros(M, 3268, device="cuda")
        for n, k, o, bt in zip(Ns, Ks, offs, Bts):
            if half:
                check(lib.rcb_gemm_tc_h(A.data_ptr() + es * o, ld, ptr(bt), k, c_one.data_ptr() + 4 * o, 3268, M, n, k,
                                        None, 1, 0, 0, stream()))
            else:
                check(lib.rcb_gemm_tc(A.data_ptr() + es * o, ld, ptr(bt), k, c_one.data_ptr() + 4 * o, 3268, M, n, k,
                                      None, 1, 0, 0, stream()))
        args, keep = FitEngine._batch_args([A.data_ptr() + es * o for o in offs], ld, Bts,
                                           [c_bat.data_ptr() + 4 * o for o in offs], 3268, M, Ns, Ks, half)
        check(lib.rcb_gemm_tc_batch(*args, stream()))
        torch.cuda.synchronize()
        assert torch.equal(c_one, c_bat)
        ref = torch.cat([A[:, o:o + k].float() @ bt.float().t() for o, k, bt in zip(offs, Ks, Bts)], 1)
        assert float((c_bat[:, :3267] - ref).abs().max()) < 5e-3 * float(ref.abs().max())
