timeout 900 python -m pytest tests -m gpu -x -q -k "lattice or eager" 2>&1 | tail -15
