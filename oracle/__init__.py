"""CPU oracle for the DPMN hot path (PGRM stack + Complementation Modulation Module).

TEST INFRASTRUCTURE ONLY.  Nothing under `dpmn_b200/` imports this package; the only legitimate
callers are `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs, and there only as the checker or the timed CPU baseline -- never as the shipped compute path.

It is a numpy restatement of the reference's PyTorch algorithm:

  oracle/pgrm_oracle.py  <- /root/reference/model/pgrm.py
  oracle/cmm_oracle.py   <- /root/reference/model/cmm.py

Parity status: PINNED.  The reference ships no golden vectors or tests, so the
fixtures under `tests/golden/` were minted by importing the unmodified reference modules in the build
container (`oracle/make_golden.py`, committed; needs /root/reference and a 3-symbol `timm` shim) and
`tests/test_oracle_golden.py` checks this restatement against them to 2e-5 (max|d| / max|ref|).
Train-mode stochastic paths (Dropout / timm DropPath RNG streams) are NOT pinned: the reference has no
test at that boundary and torch's RNG stream is not reproducible outside torch; all fixtures are eval().
"""
