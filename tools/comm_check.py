"""Two (or more) processes, one GPU each, joined by the LIBRARY's communicator (bcm3b200_comm_init) -- no torch.distributed:
rank 0 writes the communicator id to a file, the others read it (what a C++ host without MPI would do). Every rank evaluates
the same batch through bcm3b200_evaluate_batch and prints its per-chain log-likelihoods; the parent compares them with each
other (bit-identical) and with one unsharded handle.  usage: comm_check.py [world=2] [poppk|cellpop|both]"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def problems(kind):
    if kind == "poppk":
        from bcm3_b200 import synthetic as syn
        from bcm3_b200.poppk_data import PK_TWO

        prob = syn.make_poppk_problem(PK_TWO, P=501, T=8, seed=5)
        vals = syn.make_chain_values(prob, 6, seed=6)
        return prob, vals
    from bcm3_b200 import synthetic_cellpop as sc

    prob = sc.make_cellpop_problem(N=12, num_cells=203, T=12, data_cells=4, seed=11)
    return prob, sc.make_chain_values(3, seed=4)


def evaluator(kind, prob, **kw):
    if kind == "poppk":
        from bcm3_b200.poppk import PopPKEvaluator

        return PopPKEvaluator(prob, **kw)
    from bcm3_b200.cellpop import CellPopEvaluator

    return CellPopEvaluator(prob, **kw)


def child(rank, world, kind, id_path):
    from bcm3_b200 import _lib

    prob, vals = problems(kind)
    ev = evaluator(kind, prob, device=rank, shard_rank=rank, shard_count=world)
    if rank == 0:
        cid = _lib.comm_unique_id()
        with open(id_path + ".tmp", "wb") as f:
            f.write(cid)
        os.rename(id_path + ".tmp", id_path)
    else:
        t0 = time.time()
        while not os.path.exists(id_path):
            if time.time() - t0 > 120:
                raise RuntimeError("no communicator id")
            time.sleep(0.05)
        cid = open(id_path, "rb").read()
    ev.comm_init(cid)
    out = []
    for _ in range(2):  # twice: the second call reuses every buffer
        logp, status = ev.evaluate(vals)
        out.append(logp.tolist())
    ev.close()
    print("RESULT " + json.dumps({"rank": rank, "logp": out, "status": status.tolist()}), flush=True)


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    kinds = ["poppk", "cellpop"] if len(sys.argv) < 3 or sys.argv[2] == "both" else [sys.argv[2]]
    ok = True
    for kind in kinds:
        with tempfile.TemporaryDirectory() as d:
            id_path = os.path.join(d, "comm_id")
            procs = [subprocess.Popen([sys.executable, __file__, "--child", str(r), str(world), kind, id_path], stdout=subprocess.PIPE, text=True) for r in range(world)]
            results = {}
            for p in procs:
                text, _ = p.communicate(timeout=600)
                if p.returncode != 0:
                    raise RuntimeError(f"rank process failed ({p.returncode}):\n{text}")
                for line in text.splitlines():
                    if line.startswith("RESULT "):
                        r = json.loads(line[7:])
                        results[r["rank"]] = r
        prob, vals = problems(kind)
        ev = evaluator(kind, prob, device=0)
        want, _ = ev.evaluate(vals)
        ev.close()
        first = np.array(results[0]["logp"])
        same = all(np.array_equal(np.array(results[r]["logp"]), first, equal_nan=True) for r in range(world))
        rel = np.abs(first[0] - want) / np.abs(want)
        print(f"{kind}: {world} ranks bit-identical to each other and between calls: {same and np.array_equal(first[0], first[1], equal_nan=True)}; "
              f"max rel diff to the unsharded handle {rel.max():.2e}", flush=True)
        ok = ok and same and np.array_equal(first[0], first[1], equal_nan=True) and rel.max() < 1e-12
    print("COMM CHECK " + ("OK" if ok else "FAILED"))
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5])
    else:
        sys.exit(main())
