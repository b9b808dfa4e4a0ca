import os, sys
from cuda import cudart
g = int(os.environ.get("BP_L2_GRAN", "0"))
if g:
    print("set", cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity, g))
print("granularity", cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity))
sys.argv = ["tools/msm_phases_quick.py", "24"]
exec(open("tools/msm_phases_quick.py").read())
