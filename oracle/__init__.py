"""CPU oracle (TEST INFRASTRUCTURE).  Import only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs — never from the product package."""
from .binding import *  # noqa: F401,F403
