#!/bin/bash
# parameter sweep on the bench frame (spp 16): fetch threshold, leaf size, PLOC radius
for f in 4 8 12 16 24 32; do echo "fetch_min=$f"; RT_B200_FETCH_MIN=$f python scripts/profile_step.py 16 2 | tail -1; done
for lib in scripts/variants/*.so; do echo "$lib"; RT_B200_LIB=$lib python scripts/profile_step.py 16 2 | tail -1; done
