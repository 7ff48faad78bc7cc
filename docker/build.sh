#!/usr/bin/env bash
# docker/build.sh [tag]  -- builds the image from the repository root (reference: docker/build_whls.sh + Dockerfile.release)
set -euo pipefail
cd "$(dirname "$0")/.."
TAG=${1:-torchacc_b200:$(python -c "exec(open('torchacc_b200/version.py').read()); print(__version__)")}
docker build -f docker/Dockerfile -t "$TAG" .
echo "built $TAG"
