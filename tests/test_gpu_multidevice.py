"""GPU suite, boxes with >= 2 devices only: the exchange step inside the library (csrc/comm_host.cuh).
(a) one process per GPU joined by bcm3b200_comm_init -- tools/comm_check.py starts the rank processes, hands the id round
    through a file and checks that every rank's bcm3b200_evaluate_batch returns the same bits, equal to one unsharded handle;
(b) one process driving two devices through a device_count = 2 handle, both model kinds."""
import os
import subprocess
import sys

import numpy as np
import pytest

from bcm3_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_two():
    if _lib.device_count() < 2:
        pytest.skip("needs two CUDA devices")


def test_comm_unique_id_is_an_nccl_id():
    cid = _lib.comm_unique_id()
    assert len(cid) == _lib.COMM_ID_BYTES and any(cid)


def test_library_communicator_two_processes():
    _need_two()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "comm_check.py"), "2", "both"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "COMM CHECK OK" in r.stdout, r.stdout + r.stderr


def test_cellpop_two_devices_in_one_process():
    _need_two()
    from bcm3_b200 import synthetic_cellpop as sc
    from bcm3_b200.cellpop import CellPopEvaluator

    prob = sc.make_cellpop_problem(N=12, num_cells=203, T=12, data_cells=4, seed=11)
    vals = sc.make_chain_values(3, seed=4)
    one = CellPopEvaluator(prob)
    want, _ = one.evaluate(vals)
    one.close()
    two = CellPopEvaluator(prob, device_count=2)
    got, status = two.evaluate(vals)
    again, _ = two.evaluate(vals)
    two.close()
    assert (status == 0).all()
    assert np.array_equal(got, again)
    # the two-device handle sums per-shard sums (sum / count), the one-device handle divides every cell first: round-off apart
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def test_poppk_two_devices_in_one_process():
    _need_two()
    from bcm3_b200 import synthetic as syn
    from bcm3_b200.poppk import PopPKEvaluator
    from bcm3_b200.poppk_data import PK_TWO

    prob = syn.make_poppk_problem(PK_TWO, P=501, T=8, seed=5)
    vals = syn.make_chain_values(prob, 6, seed=6)
    one = PopPKEvaluator(prob)
    want, _ = one.evaluate(vals)
    one.close()
    two = PopPKEvaluator(prob, device_count=2)
    got, _ = two.evaluate(vals)
    two.close()
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
