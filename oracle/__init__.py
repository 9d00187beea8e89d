"""TEST INFRASTRUCTURE — the two parity oracles (see oracle/bindings.py).

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
bench.py may import this package; the product (klang_b200/) never does."""
from .bindings import *  # noqa: F401,F403
from .bindings import ref, port  # noqa: F401
