import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["n_gpus"], "value %.2f G" % (d["value"]/1e9), "ms/step %.2f" % d["ms_per_step"], d["config"]["fft_tile_xyz"], d["config"]["tiles_per_gpu"], d["config"]["fft_box_over_useful_voxels"], d["config"]["sharding"][:24], "| e2e %.2f G" % (d["e2e"]["value"]/1e9), d["roofline"]["all_passes_ms_per_launch"], d["config"]["output_finite"])
