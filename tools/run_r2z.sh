cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_hermitian.py tests/test_cabi_symbols.py tests/test_abi_layout.py -x -q > gpurun_out/r2z_herm.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_herm.log
