for shape in "16 256 256 96 256 1" "16 256 256 96 288 0" "16 256 256 48 128 1" "16 256 256 48 144 0"; do
  for bal in 0 1 2 3; do PIR_PWDW_BAL=$bal python tools/time_pwdw.py $shape | sed "s/\$/  bal=$bal/"; done
done
