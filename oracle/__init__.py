"""CPFFT ORACLE -- TEST INFRASTRUCTURE ONLY.  See oracle/oracle.h.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package ``cpfft_b200`` never imports this.
"""
from .binding import Oracle, build_oracle, oracle_lib_path  # noqa: F401
