"""CPU test (no GPU): the engine's pure host logic, exercised by a C++ program
built in-tree (tests/cpp/test_host_logic.cu): the staggered filter-switch pointer
table against a literal restatement of the reference's queues
(apf/apf/convolver.h:618-699) and the MAC work splitter's invariants."""
import os
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ssr-pre-0.4.0_b200", "lib")


def test_filter_switch_table_and_work_splitter():
    exe = os.path.join(LIB, "test_host_logic")
    assert os.path.exists(exe), f"{exe} missing: run __graft_entry__.build()"
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert res.returncode == 0, res.stdout
    assert "All tests passed" in res.stdout
