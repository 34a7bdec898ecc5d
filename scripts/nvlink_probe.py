"""Which NVLink traffic counters does this box expose?  Prints what NVML field values and `nvidia-smi nvlink` return
(used once to decide how bench.py measures halo traffic)."""
import subprocess

try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    for name in ("NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX", "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX", "NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX",
                 "NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_RX", "NVML_FI_DEV_NVLINK_LINK_COUNT"):
        fid = getattr(pynvml, name, None)
        print(name, fid)
        if fid is None:
            continue
        for arg in ([fid], [(fid, 0xFFFFFFFF)], [(fid, 0)]):
            try:
                v = pynvml.nvmlDeviceGetFieldValues(h, arg)
                print("   ", arg, [(x.fieldId, x.scopeId, x.nvmlReturn, x.valueType, x.value.ullVal, x.value.uiVal) for x in v])
            except Exception as e:
                print("   ", arg, "->", repr(e))
except Exception as e:
    print("pynvml:", repr(e))
for cmd in (["nvidia-smi", "nvlink", "-gt", "d", "-i", "0"], ["nvidia-smi", "nvlink", "-s", "-i", "0"], ["nvidia-smi", "topo", "-m"]):
    r = subprocess.run(cmd, capture_output=True, text=True)
    print("$", " ".join(cmd), "rc", r.returncode)
    print(r.stdout[:1500], r.stderr[:300])
