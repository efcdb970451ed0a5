#!/bin/bash
timeout 300 python scripts/sample_profile.py 2>&1 | tail -12
