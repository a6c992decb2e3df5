#!/bin/sh
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored, shipped to the GPU box by gpurun).
# /root/reference is read-only and the build writes an egg-info into the source tree: install from a copy.
# --no-deps: the only dependency is numpy, which the offline wheelhouse does not carry (it is already installed).
set -e
rm -rf /tmp/refcopy && cp -r /root/reference /tmp/refcopy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$(dirname "$0")/_ref" --upgrade /tmp/refcopy
