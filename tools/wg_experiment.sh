for cfg in "2 0 0" "2 1 0" "2 0 1" "4 0 0" "4 1 0"; do
  set -- $cfg
  FNEUS_WG_STAGES=$1 FNEUS_WG_NOCONV=$2 FNEUS_WG_CEIL=$3 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --stats-steps 0 > /tmp/o.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("/tmp/o.json"))
k=d["kernel_classes"]
print("stages=$1 noconv=$2 ceil=$3: step %.3f ms  wgrad %.3f ms  sdf_bwd %.3f  sdf_fwd %.3f  relu %.3f  elem %.3f" % (d["ms_per_step"], k["tc_wgrad_group"]["ms_per_step"], k["chain_sdf_bwd"]["ms_per_step"], k["chain_sdf_fwd"]["ms_per_step"], k["chain_relu"]["ms_per_step"], k["elementwise"]["ms_per_step"]))
PY
done
