# A/B of the pwdw variants: default (four groups where they fit) vs PIR_PWDW_NG2=1 (two groups)
for shape in "16 256 256 96 256 1" "16 256 256 96 288 0" "16 256 256 48 128 1" "16 256 256 48 144 0" "16 128 128 96 256 1" "16 128 128 96 288 0" "16 64 64 192 512 1" "16 64 64 192 576 0"; do
  python tools/time_pwdw.py $shape
  PIR_PWDW_NG2=1 python tools/time_pwdw.py $shape | sed 's/$/  (NG2)/'
done
