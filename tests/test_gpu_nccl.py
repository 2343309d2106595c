"""The NCCL path (pnp_comm.cu) against the GLOBAL oracle: spawns one process per GPU with torch.distributed.run when at
least two GPUs are visible (tests/nccl_worker.py does the checking), skips otherwise.  Run on a multi-GPU box with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`; its log is kept under profiles/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,levels", [(2, 2), (4, 1)])
def test_nccl_ranks_match_global_oracle(world, levels):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_worker.py"), str(levels)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("NCCL_WORKER_OK") == world, out.stdout[-3000:]


def _build_example(name, tmp_path):
    exe = str(tmp_path / name)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".cc"), "-L", os.path.join(ROOT, "dune_pnp_b200"), "-lpnp_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "dune_pnp_b200"), "-o", exe])
    return exe


def _driver_lines(out):
    import re
    res = []
    for line in out.splitlines():
        m = re.search(r"rank (\d+) of (\d+): (\d+) owned \+ (\d+) ghost vertices \| PNP Newton (\d+) iterations, defect (\S+) -> (\S+), (\d+) linear", line)
        if m:
            res.append(dict(rank=int(m.group(1)), world=int(m.group(2)), owned=int(m.group(3)), ghost=int(m.group(4)), its=int(m.group(5)),
                            d0=float(m.group(6)), d1=float(m.group(7)), lin=int(m.group(8))))
    return res


@pytest.mark.parametrize("world", [1, 2])
def test_cpp_driver_runs_distributed_without_python(world, tmp_path):
    """examples/stationary_pnp_distributed.cc: the reference's `mpirun -np N dune_pnp` flow on the C++ facade -- native
    partitioner (pnp_partition_build), NCCL id through a rendezvous file, nested iteration, distributed multigrid -- started
    by a process launcher with no Python around the library.  N ranks reproduce the one-rank Newton run: same iteration
    counts, same defects to 9 digits, owned vertices tile the mesh."""
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    exe = _build_example("stationary_pnp_distributed", tmp_path)
    msh = os.path.join(util.GOLDEN, "msh", "pore.msh")
    args = [util.cfg_path("pore"), msh, "3", "1"]
    runs = {}
    for w in sorted({1, world}):
        rdv = str(tmp_path / ("nccl_id_%d" % w))
        if w == 1:
            cmd = [exe] + args + [rdv]
            env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
        else:
            port = 29500 + (os.getpid() % 2000) + 7
            cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node", str(w), "--master-addr",
                   "127.0.0.1", "--master-port", str(port), exe] + args + [rdv]
            env = dict(os.environ)
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
        runs[w] = _driver_lines(out.stdout)
        assert len(runs[w]) == w, out.stdout[-2000:]
    one = runs[1][0]
    assert one["ghost"] == 0 and one["its"] >= 1 and one["d1"] < 1e-8 * one["d0"]
    if world > 1:
        assert sum(r["owned"] for r in runs[world]) == one["owned"]
        for r in runs[world]:
            assert r["ghost"] > 0 and r["its"] == one["its"]
            assert abs(r["d0"] - one["d0"]) <= 1e-9 * one["d0"]
