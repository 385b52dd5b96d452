"""The SUB sets of the reference's dist-worker integration tests (DistQoS0Test.java:82-340,450-561, transcribed in
test_oracle_golden.py) through the CUDA path: same route set as the oracle, same fan-out count as BatchDistReply reports."""
import numpy as np
import pytest

import oracle_lib as O
from test_gpu_forward import B, compare_with_oracle, make_index  # noqa: F401  (B is the module's fixture)
from test_oracle_golden import INT_MAX, OTHER_TENANT, QOS0_CASES, TENANT_ID, _qos0_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", QOS0_CASES)
def test_dist_qos0_fanout_counts_on_gpu(B, name):  # noqa: F811
    kv, topic, fanout = _qos0_case(name)
    kv.freeze()
    idx = make_index(B, sorted(kv.items()))
    tenants = [TENANT_ID, OTHER_TENANT]
    want = compare_with_oracle(B, idx, kv, tenants, [topic, topic], np.array([0, 1], np.int32), INT_MAX, INT_MAX, O.MODE_BRUTE)
    sets = want.route_sets()
    assert len(sets[0]) == fanout
    if name != "case7":
        assert len(sets[1]) == 0      # the other tenant has no routes in these cases
    idx.close()
