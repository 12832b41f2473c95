cd $GRAFT_REPO_ROOT
ncu --set full --import-source on --clock-control none -k regex:pwdwt -c 1 -o gpurun_out/r2_pwdwt_v2 -f python tools/time_pwdw.py 16 256 256 96 256 1 3 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
