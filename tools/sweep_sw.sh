export PROF_AW_K=8,16
for cfg in "4 300 30" "3 120 30" "2.5 60 25" "0 0 0"; do
  set -- $cfg
  echo "#### perp=$1 along=$2 ang=$3"
  PLSLAM_SW_PERP=$1 PLSLAM_SW_ALONG=$2 PLSLAM_SW_ANG=$3 python tools/prof_aw.py 16,640,480,synth 16,640,480,curvy 2>&1 | grep -v "kcycles" | cut -c1-330
done
