"""torchrun check of the distributed driver: every rank computes its row shard, tables are
all-gathered over NCCL, rank 0 compares with a single-GPU run and the oracle sample.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine  # noqa: E402
from tvbingefriend_recommendation_service_b200.engine import stage  # noqa: E402
from tvbingefriend_recommendation_service_b200.multi_gpu import DistributedTopK, compute_top_k_distributed  # noqa: E402
from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cat = make_catalogue(5000, 3000, nnz=30, seed=123)
    eng = HybridTopKEngine(local)
    full = compute_top_k_distributed(cat.features(), engine=eng, symmetric=False)
    sym = compute_top_k_distributed(cat.features(), engine=eng, symmetric=True)
    # end-to-end form: sliced upload + NVLink replication in, shared pinned host table out (twice: the
    # second run recycles the first run's operand buffer)
    dtk = DistributedTopK(eng, cat.n_shows, 20)
    st = stage(cat.features())
    for sym_flag in (True, False):
        e2e = dtk.run(st, symmetric=sym_flag)
        assert np.array_equal(e2e.indices, full.indices) and np.array_equal(e2e.counts, full.counts), sym_flag
        assert np.array_equal(e2e.hybrid[e2e.indices >= 0], full.hybrid[full.indices >= 0]), sym_flag
        dist.barrier()
    dtk.close()
    if dist.get_rank() == 0:
        assert np.array_equal(full.indices, sym.indices), "tile-sharded symmetric sweep: indices differ"
        assert np.array_equal(full.counts, sym.counts)
        m = full.indices >= 0
        assert np.array_equal(full.hybrid[m], sym.hybrid[m])
        single = eng.compute_top_k(cat.features())
        assert np.array_equal(full.indices, single.indices), "indices differ"
        assert np.array_equal(full.counts, single.counts), "counts differ"
        from helpers import assert_topk_matches

        rep = assert_topk_matches(full, cat.features(), np.arange(0, 5000, 53))
        print(f"distributed check ok on {dist.get_world_size()} ranks: {rep.summary()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
