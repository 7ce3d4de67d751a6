#!/bin/bash
timeout 200 python -m pytest tests -m gpu -q --timeout=150 2>&1 | tail -1 | cut -c1-200
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-300
