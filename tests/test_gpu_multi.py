"""Bit-for-bit equality of the multi-GPU run (NCCL halo exchange) with the single-GPU run; needs >= 2 GPUs on the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from transport_se_b200.advection import cuda_lib
    return cuda_lib().tse_device_count()


@pytest.mark.parametrize("nranks,ne,test,limiter", [(2, 8, 11, 8), (2, 8, 12, 8), (2, 8, 11, 0), (4, 8, 11, 8), (8, 30, 11, 8)])
def test_bit_for_bit_across_gpu_counts(nranks, ne, test, limiter):
    n = _ngpu()
    if n < nranks:
        pytest.skip("needs %d GPUs, this box has %d" % (nranks, n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + nranks + ne + limiter), os.path.join(ROOT, "tests", "mgpu_check.py"), str(ne), "5", str(test), "2", str(limiter)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0
