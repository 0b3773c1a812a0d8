python tools/ref_python_profile.py 10 > gpurun_out/ref_python_profile10.txt 2>/dev/null; head -50 gpurun_out/ref_python_profile10.txt
