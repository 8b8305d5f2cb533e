for post in class rows; do
GR_CTC_POST=$post timeout 200 python - <<'PY'
import sys, os; sys.path.insert(0, ".")
import torch, bench
print(os.environ.get("GR_CTC_POST"), bench.ctc_microbench(torch.device("cuda:0"), 6549.4)["ms"])
PY
done
GR_CTC_POST=rows timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctc_loss_grad -s 2 -c 1 -f -o gpurun_out/prof_ctc5 python scripts/micro.py ctc > gpurun_out/ncu_ctc5.log 2>&1
