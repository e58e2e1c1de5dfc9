#!/bin/bash
timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 2>&1 | tail -4
timeout 300 python scripts/diag_c3_slow.py alone
timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 --dtype f32 2>&1 | tail -4
