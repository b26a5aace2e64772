"""Per-(process, device) context of the C ABI (include/mmda_b200.h::mmda_ctx_info):
device-dependent host state is keyed by the current device, never process-global."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _info():
    from mmda_b200._lib import LIB
    out = (ctypes.c_int * 6)()
    LIB.call("mmda_ctx_info", out)
    return list(out)


def _tc_gemm(dev):
    """one 3xTF32 tcgen05 GEMM (C = A B^T from plain fp32 operands, kind 2) checked against fp64"""
    from mmda_b200.engine import Kernels, _ptr
    k = Kernels(); k.bind_stream()
    g = torch.Generator(device="cpu").manual_seed(3)
    M, N, K = 256, 384, 320
    A = torch.randn(M, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    C = torch.zeros(M, N, device=dev)
    k._c("mmda_gemm_tc", 2, 0, 0, M, N, K, _ptr(A), None, A.stride(0), _ptr(B), None, B.stride(0),
         1.0, _ptr(C), N, None, None, 0, 1, 0)
    ref = A.double() @ B.double().t()
    assert ((C.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5


def test_ctx_follows_the_current_device():
    n = torch.cuda.device_count()
    for d in range(min(n, 2)):
        with torch.cuda.device(d):
            props = torch.cuda.get_device_properties(d)
            _tc_gemm(torch.device("cuda", d))
            info = _info()
            assert info[0] == d
            assert info[1] == props.multi_processor_count
            assert info[2] >= 200 * 1024          # B200: 227 KB opt-in shared memory per block
            assert info[4] == 1                    # this device's scheduler slots exist now
            assert 0 <= info[5] <= 28672           # slots held by captured graphs of this device
    if n < 2:
        return
    # every context owns its own slots: device 1's GEMM ran on device-1 memory (the parity check
    # above would have faulted on a device-0 pointer without peer access) and the counters are
    # independent
    with torch.cuda.device(0):
        assert _info()[0] == 0
