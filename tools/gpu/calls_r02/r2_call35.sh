#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
