timeout 200 python -m pytest tests/test_herest_dropin.py -m gpu -q 2>&1 | tail -12
