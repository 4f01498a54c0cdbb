"""tcgen05 implicit-GEMM kernel against plain PyTorch fp32 references of the same op."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

import frx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["ws_tma_im2col", "ws_cp_async_gather", "one_tile_per_cta"])
def handle(request):
    cfg = frx._lib.FrxConfig()
    cfg.network, cfg.height, cfg.width, cfg.in_ch = 0, 128, 256, 1
    cfg.enc_hidden, cfg.enc_filter, cfg.enc_layers, cfg.enc_heads = 512, 512, 2, 8
    cfg.dec_src, cfg.dec_hidden, cfg.dec_filter, cfg.dec_layers, cfg.dec_heads = 512, 256, 1024, 3, 8
    cfg.num_classes, cfg.sos_id, cfg.eos_id, cfg.pad_id = 245, 0, 1, 2
    cfg.max_batch, cfg.max_steps, cfg.precision, cfg.device = 2, 4, 1, 0
    h = frx._lib.Handle(cfg)
    h.call("frx_set_option", b"tc_ws", 0 if request.param == "one_tile_per_cta" else 1)
    h.call("frx_set_option", b"tc_im2col", 1 if request.param == "ws_tma_im2col" else 0)
    return h


def _run(handle, a, w, m, n, k, conv=None, scale=None, shift=None, act=0, out_f32=True):
    c = torch.empty(m, n, device="cuda", dtype=torch.float32 if out_f32 else torch.float16)
    conv7 = (ctypes.c_int32 * 7)(*conv) if conv else None
    handle.call("frx_tc_gemm", a.data_ptr(), w.data_ptr(), c.data_ptr(), m, n, k, conv7,
                scale.data_ptr() if scale is not None else None, shift.data_ptr() if shift is not None else None,
                act, int(out_f32), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (256, 256, 256), (1000, 24, 216), (777, 160, 960),
                                   (128, 768, 128), (4096, 1536, 512), (130, 48, 96), (64, 512, 1536)])
def test_dense_gemm(handle, m, n, k):
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    a = torch.randn(m, k, device="cuda", generator=g).half()
    w = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).half()
    c = _run(handle, a, w, m, n, k)
    ref = a.float() @ w.float().t()
    assert (c - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def test_many_tiles_per_cta(handle):
    """More tiles than persistent CTAs: exercises the ring / accumulator phase wrap-around."""
    m, n, k = 128 * 700, 96, 216
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(m, k, device="cuda", generator=g).half()
    w = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).half()
    c = _run(handle, a, w, m, n, k)
    ref = a.float() @ w.float().t()
    assert (c - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def test_epilogue_scale_shift_act_bf16_out(handle):
    m, n, k = 300, 96, 216
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(m, k, device="cuda", generator=g).half()
    w = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).half()
    scale = torch.rand(n, device="cuda", generator=g) + 0.5
    shift = torch.randn(n, device="cuda", generator=g)
    c = _run(handle, a, w, m, n, k, scale=scale, shift=shift, act=2, out_f32=False)
    ref = F.silu((a.float() @ w.float().t()) * scale + shift)
    assert (c.float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("b,h,w_,cin,cout,stride", [(2, 63, 127, 24, 24, 1), (3, 63, 127, 24, 96, 2),
                                                     (2, 32, 64, 48, 192, 1), (2, 32, 64, 48, 192, 2),
                                                     (5, 16, 32, 64, 256, 1)])
def test_conv3x3_implicit_gemm(handle, b, h, w_, cin, cout, stride):
    """3x3 conv as implicit GEMM with timm's padding rule (static 1 for stride 1,
    TF-'same' (0,1) for stride 2 on even sizes / (1,1) on odd sizes)."""
    g = torch.Generator(device="cuda").manual_seed(b + cout)
    x = torch.randn(b, h, w_, cin, device="cuda", generator=g).half()          # NHWC
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (9 * cin) ** 0.5).half()
    oh, ow = -(-h // stride), -(-w_ // stride)
    wp = wt.permute(0, 2, 3, 1).contiguous().view(cout, 9 * cin)                     # [O][kh][kw][I]
    c = _run(handle, x, wp, b * oh * ow, cout, 9 * cin, conv=(b, h, w_, cin, 3, stride, 1))
    xn = x.float().permute(0, 3, 1, 2)
    if stride == 1:
        ref = F.conv2d(xn, wt.float(), None, 1, 1)
    else:
        ph = max((oh - 1) * 2 + 3 - h, 0)
        pw = max((ow - 1) * 2 + 3 - w_, 0)
        ref = F.conv2d(F.pad(xn, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]), wt.float(), None, 2, 0)
    ref = ref.permute(0, 2, 3, 1).reshape(b * oh * ow, cout)
    assert (c - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
