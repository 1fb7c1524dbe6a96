set -x
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^  \|array(\[" | cut -c1-400 | tail -25
