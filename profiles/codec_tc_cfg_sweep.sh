for cfg in "1,2" "1,1" "2,1" "2,2"; do echo "== cfg $cfg"; LADIFF_CODEC_TC_CFG=$cfg timeout 120 python profiles/codec_prof.py 2 2>&1 | awk '/pass 1/{p=1} p' | grep "conv" | awk '{s+=$2; printf "%s ", $2} END {print " | sum", s}'; done
echo "== default"; timeout 120 python profiles/codec_prof.py 2 2>&1 | awk '/pass 1/{p=1} p' | grep "conv" | awk '{s+=$2; printf "%s ", $2} END {print " | sum", s}'
