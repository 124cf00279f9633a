timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "Error|error|assert|FAILED" | head -8
