"""CPU oracle (TEST INFRASTRUCTURE ONLY).

Restatement of the reference's algorithm for the waveform -> log-mel -> Conformer -> CTC hot path, in
numpy / plain torch-CPU float64/float32.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product package never does.
Parity pinning: the reference ships no tests or golden vectors; every function here is
pinned by running the reference itself in the authoring container (tests/test_oracle_vs_reference.py,
tools/make_golden.py) and by the committed fixtures under tests/golden/.
"""
