for v in g2b5 g2b6 base; do cp tools/alt/$v.so vector_db_id_compression_b200/libidcodec.so; echo $v; bash tools/ab.sh "V=$v" --zipf-s 0 2>&1 | tail -1; done
cp tools/alt/g2b5.so vector_db_id_compression_b200/libidcodec.so; bash tools/ab.sh "V=g2b5" 2>&1 | tail -1
