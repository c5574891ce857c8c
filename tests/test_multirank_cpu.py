"""world_size-2 gloo run of the bench's N > 1 path on CPU: rank 0 generates the shared inputs, the
other rank waits for them, timing is the max over ranks and the value the whole-job aggregate."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_gloo(tmp_path):
    env = dict(os.environ, MR_BENCH_BACKEND="gloo", MR_BENCH_DIR=str(tmp_path), OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531",
                        os.path.join(ROOT, "tests", "helpers", "multirank_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    text = r.stdout.decode().replace("}{", "}\n{")           # two ranks, one pipe: lines may still run together
    rows = [json.loads(l) for l in text.splitlines() if l.startswith("{")]
    assert sorted(x["rank"] for x in rows) == [0, 1]
    assert all(x["files_exist"] for x in rows)
    assert rows[0]["reads"] == rows[1]["reads"] > 0
    # rank r aligns shard r: disjoint reads of the same genome, numbered on from the previous shard
    rows.sort(key=lambda x: x["rank"])
    assert [x["shard"] for x in rows] == [0, 1] and rows[0]["reads_file"] != rows[1]["reads_file"]
    assert rows[0]["first_read"].startswith("read0/") and rows[1]["first_read"].startswith("read%d/" % rows[0]["reads"])
    for x in rows:
        assert x["t_max"] == 2.0                       # slowest rank
        assert x["total"] == 2 * x["reads"]
        assert x["value"] == 2 * x["reads"] / 2.0


def test_reference_arm_other_ranks_do_no_work(tmp_path):
    """bench.py --impl reference: only rank 0 runs and prints; every other rank exits 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", MR_BENCH_DIR=str(tmp_path))
    # rank 1 would wait for rank 0's data: create the done marker by running rank 0's generation first
    env0 = dict(env, RANK="0", LOCAL_RANK="0")
    code = ("import sys; sys.path.insert(0, %r); import bench, argparse; "
            "bench.data_files(argparse.Namespace(genome=40000, coverage=2.0, gpus=2, impl='reference'))" % ROOT)
    subprocess.check_call([sys.executable, "-c", code], env=env0)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--genome", "40000", "--coverage", "2", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == b""
