mkdir -p gpurun_out/r03
timeout 400 python profiles/probe_pipeline.py > gpurun_out/r03/probe_pipeline.json 2> gpurun_out/r03/probe_pipeline.err; cat gpurun_out/r03/probe_pipeline.json; tail -5 gpurun_out/r03/probe_pipeline.err
