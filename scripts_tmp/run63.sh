timeout 600 python profiles/r2_capacity_tier.py > gpurun_out/r2f_capacity_tier.json 2> gpurun_out/r2f_capacity_tier.err; cat gpurun_out/r2f_capacity_tier.json | cut -c1-700
timeout 900 python profiles/r2_batch_pipeline.py > gpurun_out/r2f_batch_pipeline.json 2> gpurun_out/r2f_batch_pipeline.err; cat gpurun_out/r2f_batch_pipeline.json | tr -d '\n' | cut -c1-1500
