#!/bin/bash
timeout 300 python scripts/profile_c4.py --reps 10 2>&1 | cut -c1-40
