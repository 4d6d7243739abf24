"""oracle/ — CPU restatement of the reference's algorithms for the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under go-slam_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may, and there
only as the checker / the timed CPU baseline, never as a fallback for the CUDA path.
See each module's header for the reference file:line it follows and how it is pinned.
"""
