import sys, json
sys.path.insert(0, ".")
import bench
print(json.dumps(bench.matching_methods_extras(0), indent=0)[:1800])
