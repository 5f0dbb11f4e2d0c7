"""`python -m zeekstd_b200 ...` = the reference's `zeekstd` binary (cli/src/main.rs:12-31) over the GPU codec"""
import sys

from .cli import main

sys.exit(main())
