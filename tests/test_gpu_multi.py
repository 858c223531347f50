"""Multi-GPU correctness of the view-sharded exchange (-m gpu; needs >= 2 GPUs, skipped on a single-GPU box):
2 ranks x V views through ``ViewShardedExchange`` + ``DensificationStats.sync()`` == 1 process x 2V views, identical sums
on every rank, and identical Gaussian sets after ``refine``.  The work is done by ``tests/multi/worker_exchange.py``
under ``torch.distributed.run``."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(world, views, mc=True, port=29631, solo=False):
    env = dict(os.environ, FG_TEST_VIEWS_PER_RANK=str(views), FG_XCHG_NO_MULTICAST="" if mc else "1",
               FG_XCHG_FORCE_MULTICAST="1" if mc else "", FG_XCHG_SOLO="1" if solo else "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "multi" / "worker_exchange.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0 and "MULTI-OK" in out.stdout, out.stdout[-3000:] + out.stderr[-6000:]
    return out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("views", [1, 2, 5])  # 5 views x 2 ranks: more than one eight-view chunk of fg_xchg_sh_bwd_views
def test_two_ranks_equal_one_process(built_lib, views):
    _run(2, views, port=29631 + views)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_without_multicast(built_lib):
    """The peer load / store form of the all-reduce kernel (the default below 4 ranks, and without an NVSwitch multicast
    object); the two tests above force the in-switch (multimem) form."""
    _run(2, 1, mc=False, port=29641)


def test_single_rank_group_runs_the_exchange_kernels(built_lib):
    """One GPU is enough to execute the exchange path itself: a one-rank group with FG_XCHG_SOLO publishes the colour gradients
    of 9 views (two eight-view chunks), rebuilds the SH rows with fg_xchg_sh_bwd_views and passes the geometry through the
    all-reduce kernel; the result must equal the plain backward over the same 9 views."""
    _run(1, 9, mc=False, port=29651, solo=True)
