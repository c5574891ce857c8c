"""Run under torchrun with MR_BENCH_BACKEND=gloo: exercises bench.py's multi-rank plumbing on CPU
(shared synthetic inputs generated once by rank 0, barrier, max / sum over ranks, weak-scaling
aggregation).  Each rank prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

args = argparse.Namespace(gpus=int(os.environ["WORLD_SIZE"]), genome=40000, coverage=2.0)
w, files = bench.data_files(args)
dist = bench.dist_setup(args.gpus)
rank = int(os.environ["RANK"])
bench.barrier(dist)
names = [line[1:].split()[0] for line in open(files["reads"]) if line.startswith(">")]
reads = len(names)
t = 1.0 + rank                                  # pretend rank r needed 1 + r seconds
t_max = bench.max_over_ranks(dist, t)
total = bench.sum_over_ranks(dist, float(reads))
value = args.gpus * reads / t_max               # whole-job throughput = all ranks' units / slowest rank
# one write per line: the two ranks share the launcher's stdout pipe
sys.stdout.write(json.dumps({"rank": rank, "reads": reads, "shard": files["shard"], "reads_file": files["reads"],
                             "first_read": names[0], "last_read": names[-1], "t_max": t_max, "total": total, "value": value,
                             "files_exist": all(os.path.exists(files[k]) for k in ("sr", "reads", "unitigs"))}) + "\n")
sys.stdout.flush()
dist.destroy_process_group()
