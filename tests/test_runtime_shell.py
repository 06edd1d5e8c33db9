"""CPU tests (-m "not gpu") of the `CRFP_runtime.MRCF_simple_v18` drop-in: the oracle
restatement against golden outputs of the REAL reference class (oracle/make_golden_runtime.py runs it on the CPU behind
import shims), and the product shell's wiring through the test kernel set.  GPU twin: tests/test_gpu_zzz_runtime.py."""
import os

import pytest
import torch

import hostemu
from crfp_b200.runtime import MRCF_simple_v18, make_runtime_state_dict, runtime_param_shapes
from crfp_b200.synthetic import make_clip
from oracle import crfp_oracle as O

CASES = ["runtime_full_n1_t3_16x24", "runtime_region_n2_t3_16x24"]


def _case(golden_dir, name):
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    c = fix["case"]
    sd = make_runtime_state_dict(seed=c["wseed"])
    assert abs(float(sum(v.double().sum() for v in sd.values())) - fix["weights_sum"]) < 1e-6
    lrs, _, _, _ = make_clip(seed=c["seed"], n=c["n"], t=c["t"], h=c["h"], w=c["w"], fv_size=16)
    fvs = torch.rand(c["n"], c["t"], 3, c["fh"], c["fw"], generator=torch.Generator().manual_seed(c["fv_seed"]))
    return fix, sd, lrs, fvs, tuple(c["warp"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name, golden_dir):
    fix, sd, lrs, fvs, warp = _case(golden_dir, name)
    out = O.runtime_v18_forward(sd, lrs, fvs, warp)
    assert out.shape == fix["out"].shape and (out - fix["out"]).abs().max().item() <= 1e-5


@pytest.mark.parametrize("name", CASES)
def test_shell_matches_reference_golden(name, golden_dir):
    fix, sd, lrs, fvs, warp = _case(golden_dir, name)
    model = MRCF_simple_v18("cpu", mid_channels=32, kernels=hostemu.HostEmuKernelSet())
    assert len(model.state_dict()) == 158 and set(model.state_dict().keys()) == set(runtime_param_shapes().keys())
    model.load_state_dict(sd, strict=True)
    out = model(lrs, fvs, warp_size=warp)
    assert out.shape == fix["out"].shape and (out - fix["out"]).abs().max().item() < 1e-4


def test_shell_argument_errors():
    from crfp_b200 import _lib
    with pytest.raises(_lib.CrfpError):
        MRCF_simple_v18("cpu", mid_channels=16)
    model = MRCF_simple_v18("cpu", mid_channels=32)
    with pytest.raises(ValueError):
        model(torch.rand(1, 2, 3, 8, 8), torch.rand(1, 2, 3, 128, 128))          # fovea larger than the HR frame
    with pytest.raises(_lib.CrfpError):
        model(torch.rand(1, 2, 3, 8, 8), torch.rand(1, 2, 3, 32, 32))            # CPU tensors, CUDA kernel set: no fallback
    with pytest.raises(TypeError):
        model.init_weights(pretrained=3)
