timeout 900 python -m pytest tests -m gpu -q -k "qualifier or static_features" 2>&1 | tail -15
