"""GPU: stage-2 flow training kernels (csrc/flow_bwd.cu, csrc/train_enc.cu, glare_b200/flow_train.py, encoder_train.py; BASELINE config 4).

The checks run in CHILD processes (a faulting kernel must not poison the CUDA context of the rest of the suite) and are hard asserts: the
kernels have a recorded green hardware run (profiles/r41_train_check.log)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _child(script, env=None, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(HERE, script)], capture_output=True, text=True, timeout=timeout, env=env)
    log = (r.stdout + "\n" + r.stderr)[-6000:]
    print(log)
    return r.returncode, log


def test_flow_training_kernels_in_child_process(glare_lib):
    """every kernel of flow_bwd.cu against the torch restatement of its contract, the flow step against the CPU specification
    (oracle/flow_backward.py), the whole stage-2 step (encoder tape + flow objective) against the same host logic on torch primitives"""
    rc, log = _child("flow_train_gpu_check.py")
    assert rc == 0, log


def test_stage3_training_path_in_child_process(glare_lib):
    """stage-3 (deformable decoder + losses): MS-SSIM kernels and the VGG perceptual loss against the oracle, one whole evaluation through
    the drop-in modules against the unmodified reference's gradients (tests/golden/stage3.npz)"""
    rc, log = _child("stage3_gpu_check.py")
    assert rc == 0, log


def test_conv_default_kernel_in_child_process(glare_lib):
    """the shipped conv kernels (two-ring patch staging, RING2, for the 256-wide N tiles since round 2) against cuDNN fp32 on three shapes"""
    env = dict(os.environ)
    env.pop("GLARE_CONV_NO_RING2", None)
    rc, log = _child("conv_ring2_gpu_check.py", env)
    assert rc == 0, log


def test_conv_one_ring_kernel_in_child_process(glare_lib):
    """the one-ring kernel for the 256-wide tiles (GLARE_CONV_NO_RING2=1, the A/B switch; round 1's default) stays correct"""
    env = dict(os.environ)
    env["GLARE_CONV_NO_RING2"] = "1"
    rc, log = _child("conv_ring2_gpu_check.py", env)
    assert rc == 0, log
