"""CPU: simxns_b200.losses.select_layer against the reference's own LEAD/util.py:206-254 run over every branch
(tests/golden/lead_select_layer.json, oracle/make_golden.py select_layer): index work, bit-exact, same `random` stream."""
import json
import random
import types

from simxns_b200.losses import select_layer


def test_select_layer_matches_the_reference(golden_dir):
    cases = json.loads((golden_dir / "lead_select_layer.json").read_text())
    assert len(cases) == 48
    for c in cases:
        random.seed(c["seed"])
        db, teacher = select_layer(types.SimpleNamespace(**c["args"]))
        assert db == c["db"] and teacher == c["teacher"], c["args"]
