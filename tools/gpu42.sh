timeout 900 python -m pytest tests -x -q -m gpu --timeout 200 2>&1 | tail -5
