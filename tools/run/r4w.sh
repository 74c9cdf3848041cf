cp tools/alt/localA.so vector_db_id_compression_b200/libidcodec.so
python -m pytest tests/test_gpu_parity.py -x -q -k "roc" 2>&1 | tail -2
bash tools/ab.sh "LOCAL_A=1" 2>&1 | tail -1
bash tools/ab.sh "LOCAL_A=1" --zipf-s 0 2>&1 | tail -1
cp tools/alt/base.so vector_db_id_compression_b200/libidcodec.so
bash tools/ab.sh "LOCAL_A=0" 2>&1 | tail -1
