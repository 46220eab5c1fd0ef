python -m pytest tests -m gpu -q 2>&1 | tail -6
python __graft_entry__.py smoke 2>&1 | tail -4
