"""Our state_dict schema == the reference's own state_dict() key/shape dump (tests/golden/reference_state_dict_schema.json)."""
import json
import os

from dpmn_b200.schema import PGRMConfig, cmm_schema, pgrm_schema
from tests.util import GOLDEN


def _dump():
    with open(os.path.join(GOLDEN, "reference_state_dict_schema.json")) as f:
        return json.load(f)


def test_pgrm_schema_matches_reference():
    d = _dump()
    for it, mode in ((0, False), (2, False), (5, True)):
        ref = {k: tuple(v) for k, v in d[f"pgrm_iter{it}_mode{int(mode)}"].items()}
        ours = {n: tuple(s) for n, s, _ in pgrm_schema(PGRMConfig(iter=it, mode=mode))}
        assert ours == ref
        nparam = sum(int(__import__("numpy").prod(s)) for n, s, k in pgrm_schema(PGRMConfig(iter=it, mode=mode)) if k == "param")
        assert nparam == d[f"pgrm_iter{it}_mode{int(mode)}_nparams"]


def test_cmm_schema_matches_reference():
    d = _dump()
    ref = {k: tuple(v) for k, v in d["cmm_cnum64"].items()}
    ours = {n: tuple(s) for n, s, _ in cmm_schema(3, 64)}
    assert ours == ref
    import numpy as np
    assert sum(int(np.prod(s)) for n, s, k in cmm_schema(3, 64) if k == "param") == d["cmm_cnum64_nparams"] == 53583683
