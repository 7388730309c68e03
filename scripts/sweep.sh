#!/bin/bash
python scripts/profile_step.py 64 2 | tail -1
for lib in scripts/variants/*.so; do echo "$lib"; RT_B200_LIB=$lib python scripts/profile_step.py 64 2 | tail -1; done
