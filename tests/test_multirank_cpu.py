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


def test_ranks_are_pinned_to_disjoint_cores_and_subprocess_arms_are_not():
    """bench.pin_rank: every local rank keeps to its own share of the cores (all of them together cover the box once);
    the subprocess arms (reference binary, drop-in binary) are started through taskset with the whole core set."""
    code = ("import sys, os, json; sys.path.insert(0, %r); import bench\n"
            "r, n = int(sys.argv[1]), int(sys.argv[2])\n"
            "before = sorted(os.sched_getaffinity(0))\n"
            "info = bench.pin_rank(r, n)\n"
            "print(json.dumps({'before': before, 'after': sorted(os.sched_getaffinity(0)), 'info': info, 'cmd': bench.unpin(['true'])}))\n" % ROOT)
    env = dict(os.environ)
    env.pop("MR_BENCH_PIN", None)
    rows = [json.loads(subprocess.run([sys.executable, "-c", code, str(r), "2"], stdout=subprocess.PIPE, check=True, env=env).stdout) for r in range(2)]
    every = rows[0]["before"]
    if len(every) < 2:
        return
    assert sorted(rows[0]["after"] + rows[1]["after"]) == every and not set(rows[0]["after"]) & set(rows[1]["after"])
    for x in rows:
        assert x["info"] is not None
        cmd = x["cmd"]
        assert cmd[-1] == "true" and (len(cmd) == 1 or (cmd[0].endswith("taskset") and cmd[2] == ",".join(map(str, every))))
    one = json.loads(subprocess.run([sys.executable, "-c", code, "0", "1"], stdout=subprocess.PIPE, check=True, env=env).stdout)
    assert one["info"] is None and one["after"] == one["before"] and one["cmd"] == ["true"]
