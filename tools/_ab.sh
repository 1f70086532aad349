B=/root/repo/build
V=""
for n in p1024n32 p896n32 p768n32 p640n32m2; do V="$V -- ERTB_LIB=$B/libertb_$n.so"; done
python tools/ab_knobs.py --config c5 --spp-log2 24 -- "" $V 2>&1 | cut -c1-330
python tools/ab_knobs.py --config polarized_aerosol_tab_pp --spp-log2 20 -- "" $V 2>&1 | cut -c1-330
for cfg in "c4 16 0" "c4 11 1"; do set -- $cfg; python tools/ab_knobs.py --config $1 --spp-log2 $2 --sensor $3 -- "" 2>&1 | cut -c1-200; done
