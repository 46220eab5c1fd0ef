import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import bench
w = bench.build_workload("cfg2", 1234)
w["sym"].tofile("gpurun_out/cfg2_sym.bin")
print(w["sym"].shape, w["sym"].dtype, (w["sym"]==0).mean())
