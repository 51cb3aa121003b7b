#!/bin/bash
# Install the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with the snapshot).
# Used by tests/test_reference_boundary_gpu.py: the reference's own build_foundation_model / OpSlot binding with b200 ops.
set -e
cd "$(dirname "$0")/.."
rm -rf baseline/_ref
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /root/reference
