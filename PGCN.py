#!/usr/bin/env python
"""Drop-in launcher with the reference's name and flags (GPU/PGCN.py):
    python PGCN.py -a A.mtx -p A.mtx.<k>.<hp|gp|rp> -b nccl -s <k> -l <layers> -f <features>
One process per GPU; rank/size from SLURM_PROCID/SLURM_NPROCS or RANK/WORLD_SIZE (torchrun)."""
import sys

import pgcn_b200  # noqa: F401  (import shim for the hyphenated package directory)
from pgcn_b200.pgcn import main

if __name__ == "__main__":
    main(sys.argv[1:])
