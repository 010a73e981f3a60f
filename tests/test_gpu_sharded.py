"""Sharded-state path on the GPU: CudaEngine (libfhsim through the C-ABI) under 2 and 4 ranks.

The driver's `-m gpu` box has one GPU, and NCCL refuses two ranks on one device, so here the ranks share
cuda:0 and exchange slabs through the host with gloo; every kernel involved (lowered pair/diag ops through the
tile scheduler, fh_state_swap_bits, K2 write/accumulate on per-layout tables, K3 on per-layout entry lists)
is the production one.  The NCCL/NVLink exchange itself is exercised by tools/bench_sharded.py under torchrun
(`gpurun --gpus 2/4/8`), whose output is committed under profiles/.
"""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from test_sharded_host import _free_port, build_case, run_case  # noqa: E402

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, spec):
    import torch.distributed as dist
    sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "quantum-simulation-of-fermi-hubbard-model_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fhsim.sharded import CudaEngine
        g = world.bit_length() - 1
        case = build_case(*spec)
        sim = run_case(lambda n: CudaEngine(n - g, 0, dist), case)
        assert sim.swap_count > 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,spec", [
    (2, (2, 2, 4.0, 5, 11, False)),
    (2, (2, 3, 4.0, 6, 12, True)),
    (4, (2, 3, 4.0, 4, 13, False)),
    (2, (3, 3, 6.0, 6, 14, True)),
])
def test_sharded_cuda_engine_matches_oracle(world, spec):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), spec), nprocs=world, join=True)


def test_single_rank_cuda_engine():
    from fhsim.sharded import CudaEngine
    case = build_case(2, 3, 4.0, 5, 3, True)
    sim = run_case(lambda n: CudaEngine(n, 0, None), case)
    assert sim.swap_count == 0
