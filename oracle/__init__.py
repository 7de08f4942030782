"""CPU oracle for the PoseDiffusion sampling hot path -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs.  The product package (`posediffusion_b200`) never imports it.
"""
