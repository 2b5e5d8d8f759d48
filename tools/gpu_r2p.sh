set -x
timeout 600 python tools/quickstart_check.py 2>&1 | grep -v "WARNING\|^###\|st/s" | tail -12
