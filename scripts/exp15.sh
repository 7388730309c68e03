#!/bin/bash
for mb in 5 8; do RT_B200_LIB=scripts/variants/librt_lmb$mb.so python scripts/sweep2.py c2 12:16; done
