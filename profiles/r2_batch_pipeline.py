"""profiles/r1_batch_pipeline.py re-run with the round-2 kernel: many locus pairs with few samples each, wall-clock throughput of
coalgpu_run_sampler_batch, host table construction included. Output: profiles/r2_batch_pipeline.json."""
import runpy, os
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r1_batch_pipeline.py"), run_name="__main__")
