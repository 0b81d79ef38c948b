"""CPU oracle for the BlockMatching hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (gpu_stereo_matching_b200) never does.
"""
