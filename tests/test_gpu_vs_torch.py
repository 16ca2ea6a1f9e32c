"""Numerics of the sm_100a kernels against a plain PyTorch fp32 reference of the same op.

The reference below is the numerics contract spelled out with torch tensor ops: zero padding, nine
taps in row-major order, each product and each partial sum rounded to fp32 separately (torch's
elementwise mul and add are separate kernels, so nothing is contracted into an FMA), truncation
toward zero, saturation to [0, 255]. Every arithmetic path of the CUDA kernels must match it byte for
byte — for the integer SWAR paths that is the exactness argument of DESIGN.md §2 put to the test.
"""
import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter

pytestmark = pytest.mark.gpu


def torch_reference(img: np.ndarray, taps: np.ndarray, loops: int, device: str) -> np.ndarray:
    import torch
    import torch.nn.functional as F
    x = torch.from_numpy(img.astype(np.float32)).to(device)
    if x.ndim == 2:
        x = x[:, :, None]
    H, W, _ = x.shape
    h = torch.tensor(taps, dtype=torch.float32, device=device)
    for _ in range(loops):
        p = F.pad(x.permute(2, 0, 1), (1, 1, 1, 1)).permute(1, 2, 0)    # zero padding
        val = torch.zeros_like(x)
        for k in range(3):
            for l in range(3):
                prod = p[k:k + H, l:l + W, :] * h[k, l]                   # rounded product
                val = val + prod                                          # rounded sum, tap order
        x = torch.clamp(torch.trunc(val), 0.0, 255.0)
    return x.to(torch.uint8).cpu().numpy().reshape(img.shape)


@pytest.mark.parametrize("name", ["gaussian", "box", "edge", "sharpen", "laplacian", "emboss", "gaussian_soft"])
@pytest.mark.parametrize("c", [1, 3])
@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_kernels_match_plain_torch_fp32(rng, name, c, device):
    from parallel_convolution_b200 import ConvEngine
    img = random_image(rng, 150, 210, c)
    f = Filter.builtin(name)
    ref = torch_reference(img, f.taps, 4, device)
    with ConvEngine(width=210, height=150, channels=c, filter=f, backend="fused") as eng:
        out = eng.run(img, 4)
    assert np.array_equal(out, ref), f"{name} c={c}: {(out != ref).sum()} bytes differ from torch fp32 ({device})"


def test_virtual_ranks_match_torch(rng):
    from parallel_convolution_b200 import ConvEngine
    img = random_image(rng, 192, 256, 3)
    f = Filter.builtin("edge")
    ref = torch_reference(img, f.taps, 5, "cuda")
    for mode in ("stream", "resident"):
        with ConvEngine(width=256, height=192, channels=3, filter=f, backend="fused", workers=4, mode=mode) as eng:
            assert np.array_equal(eng.run(img, 5), ref), mode


def test_run_tensor_device_to_device(rng):
    """ConvEngine.run_tensor: uint8 CUDA tensor in, uint8 CUDA tensor out, no host staging — on one tile
    and on 4 virtual ranks, against the plain torch fp32 reference."""
    import torch
    from parallel_convolution_b200 import ConvEngine
    for c, name in ((3, "gaussian"), (1, "sharpen")):
        img = random_image(rng, 200, 264, c)
        f = Filter.builtin(name)
        ref = torch_reference(img, f.taps, 6, "cuda")
        x = torch.from_numpy(img).cuda()
        for workers in (1, 4):
            with ConvEngine(width=264, height=200, channels=c, filter=f, backend="fused", workers=workers) as eng:
                y = eng.run_tensor(x, 6)
                assert y.is_cuda and y.dtype == torch.uint8 and y.shape == x.shape
                assert np.array_equal(y.cpu().numpy(), ref), (c, name, workers)
                y2 = eng.run_tensor(y, 0)                 # zero iterations: a copy
                assert torch.equal(y2, y)
                with pytest.raises(ValueError):
                    eng.run_tensor(x.cpu(), 1)
