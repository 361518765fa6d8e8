#!/bin/bash
# compute-sanitizer passes on small universes (run on a GPU box).  memcheck must report 0 errors; racecheck reports the
# by-design same-value writes of the walk's stacks (all lanes of a group store the same entry) as warnings, no errors.
set -e
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
    -k "small_and_ragged or reference_test_universes or coincident or copy_vertices or theta_sweep or vote_width_32 or very_deep or deep_walk or vertex or physical_order or async_upload or reupload or pipelined_host_loop or upload_from_device"
compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "small_and_ragged or very_deep" | tail -8
