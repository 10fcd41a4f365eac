"""CPU oracle of the beat_this inference path: TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference)."""
