"""Which NVLink counters does this box expose?  (diagnostic; prints what it finds)"""
import subprocess
import pynvml as nv

nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
for scope in (0xFFFFFFFF, 0, 1, 17):
    try:
        vals = nv.nvmlDeviceGetFieldValues(h, [(nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, scope), (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, scope),
                                               (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_RX, scope), (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX, scope)])
        print("scope", scope, [(v.fieldId, v.nvmlReturn, int(v.value.ullVal)) for v in vals])
    except Exception as e:
        print("scope", scope, "error", e)
for name in ("nvmlDeviceGetNvLinkState", "nvmlDeviceGetNvLinkVersion"):
    try:
        print(name, [getattr(nv, name)(h, l) for l in range(18)])
    except Exception as e:
        print(name, "error", e)
for cmd in (["nvidia-smi", "nvlink", "-gt", "d", "-i", "0"], ["nvidia-smi", "nvlink", "-s", "-i", "0"], ["nvidia-smi", "topo", "-m"]):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print("$", " ".join(cmd), "->", r.returncode)
    print(r.stdout[:1500])
