#!/bin/bash
# round-2 probe 41: the final binary -- GPU suite and smoke
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
