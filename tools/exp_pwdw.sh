for dbg in 0 7 15 23 31 8; do PIR_PWDW_DBG=$dbg python tools/time_pwdw.py 16 256 256 96 288 0; done
for dbg in 0 7 15 31 8; do PIR_PWDW_DBG=$dbg python tools/time_pwdw.py 16 256 256 96 256 1; done
