for v in "" _b2 _b4; do
  IMMOCO_LIB_PATH=$PWD/miccai24_immoco_b200/libimmoco_b200$v.so timeout 300 python tools/hg_levels.py > gpurun_out/r87_hg$v.log 2>&1
  IMMOCO_LIB_PATH=$PWD/miccai24_immoco_b200/libimmoco_b200$v.so timeout 300 python tools/pdl_ab.py > gpurun_out/r87_iter$v.log 2>&1
done
timeout 300 python tools/fused_ab.py > gpurun_out/r87_fused.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r87_pytest.log 2>&1
