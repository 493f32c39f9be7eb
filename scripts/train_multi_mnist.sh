#!/usr/bin/env bash
# counterpart of the reference's scripts/train_multi_mnist.sh
cd "$(dirname "$0")/.." && python scripts/train_multi_mnist.py "$@"
