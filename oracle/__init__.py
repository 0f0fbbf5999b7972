"""CPU oracle for the EDM sampling / denoising hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, in plain functional torch-CPU arithmetic (fp32 or fp64), the algorithm the reference
(AgentCooper2002/AudioDiffuser) implements for the path named in BASELINE.json: EDM
preconditioning, the Heun/Euler samplers, the Karras schedule, the DSM loss and the DiffWave
residual stack. Every function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import it, and only as the checker or the timed CPU baseline. The product package
(`audiodiffuser_b200`) never imports it and raises if its CUDA library is missing.

Parity pin: the reference ships no golden vectors or numeric tests for this path,
so the oracle is pinned against outputs of the reference's own Python code run in the build
container: `oracle/make_golden.py` imports `/root/reference`, feeds it seeded inputs and writes
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks the oracle against those files.
"""
