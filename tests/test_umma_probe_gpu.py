"""tcgen05 / TMEM building-block self-test (csrc/dev/umma_probe.cu, in the dev-probe library libb200probe.so)."""
import pytest
import torch

from implicit_depth_b200 import _abi

pytestmark = pytest.mark.gpu


def run_probe(A, Bm, mode):
    D = torch.full((128, Bm.shape[0]), float("nan"), device="cuda")
    _abi.call_dev("b200_umma_probe", _abi.ptr(A), _abi.ptr(Bm), _abi.ptr(D), A.shape[1], Bm.shape[0], mode,
              _abi.stream_ptr())
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("K,N", [(64, 128), (192, 128), (128, 64), (64, 16)])
def test_bf16_mma_matches_host(mode, K, N):
    torch.manual_seed(K + N + mode)
    A = torch.randn(128, K, device="cuda")
    Bm = torch.randn(N, K, device="cuda")
    D = run_probe(A, Bm, mode)
    ref = (A.bfloat16().double() @ Bm.bfloat16().double().t()).float()
    err = (D - ref).abs().max().item()
    assert err < 1e-3 * ref.abs().max().item(), f"mode {mode} K={K} N={N}: max err {err}"


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("K,N", [(64, 128), (192, 128)])
def test_split_bf16_is_fp32_grade(mode, K, N):
    torch.manual_seed(7 * K + mode)
    A = torch.randn(128, K, device="cuda")
    Bm = torch.randn(N, K, device="cuda")
    D = run_probe(A, Bm, mode)
    ref = (A.double() @ Bm.double().t())
    rel = ((D.double() - ref).abs().max() / ref.abs().max()).item()
    assert rel < 2e-5, f"mode {mode}: split-bf16 relative error {rel}"


def test_halo_patch_descriptor():
    """Shifted start address (11 rows) + 1280-byte group stride inside a swizzled patch."""
    torch.manual_seed(5)
    A = torch.randn(128, 64, device="cuda")
    Bm = torch.randn(128, 64, device="cuda")
    D = run_probe(A, Bm, 4)
    ref = (A.bfloat16().double() @ Bm.bfloat16().double().t()).float()
    err = (D - ref).abs().max().item()
    assert err < 1e-3 * ref.abs().max().item(), f"halo descriptor: max err {err}"


@pytest.mark.parametrize("N", [128, 64, 32])
def test_sw64_halo_patch_descriptor(N):
    """64-byte swizzle, 32-channel K chunk, A rows inside a halo patch (start shifted by 11 pixel records of
    64 B, 640-byte group stride): the operand addressing of the 32-channel-block conv kernel."""
    torch.manual_seed(9 + N)
    A = torch.randn(128, 32, device="cuda")
    Bm = torch.randn(N, 32, device="cuda")
    D = run_probe(A, Bm, 5)
    ref = (A.bfloat16().double() @ Bm.bfloat16().double().t()).float()
    err = (D - ref).abs().max().item()
    assert err < 1e-3 * ref.abs().max().item(), f"sw64 halo descriptor: max err {err}"
