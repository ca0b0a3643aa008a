#!/bin/bash
# exactly what the driver runs at round end
timeout -s KILL 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
