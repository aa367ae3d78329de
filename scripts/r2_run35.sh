cd $GRAFT_REPO_ROOT/ab_tmp/ref
timeout 900 python -m pytest tests -q -p no:cacheprovider -x --co -q 2>&1 | tail -3
timeout 900 python -m pytest tests -q -p no:cacheprovider 2>&1 | tail -40 | cut -c1-250
