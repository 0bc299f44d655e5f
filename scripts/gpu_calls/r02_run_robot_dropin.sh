set -x
REF=baseline/_ref/reart
python scripts/run_reference_dropin.py --summary gpurun_out/r02_dropin_kin.json -- --seq_path={REF}/demo_data/data/nao --save_root=gpurun_out/dropin_kin --cano_idx=2 --evaluate --model=kinematic --resume={REF}/demo_data/pretrained/nao/kinematic-2/model.pth.tar > gpurun_out/r02_dropin_kin.log 2>&1
tail -3 gpurun_out/r02_dropin_kin.log
python scripts/run_reference_dropin.py --summary gpurun_out/r02_dropin_base.json -- --seq_path={REF}/demo_data/data/nao --save_root=gpurun_out/dropin_base --cano_idx=2 --evaluate --model=base --resume={REF}/demo_data/pretrained/nao/base-2/model.pth.tar > gpurun_out/r02_dropin_base.log 2>&1
tail -3 gpurun_out/r02_dropin_base.log
timeout 900 python scripts/run_reference_dropin.py --summary gpurun_out/r02_dropin_train.json -- --seq_path={REF}/demo_data/data/nao --save_root=gpurun_out/dropin_train --cano_idx=2 --model=base --n_iter=200 --assign_iter=100 --use_assign_loss --snapshot_gap=100 > gpurun_out/r02_dropin_train.log 2>&1
tail -3 gpurun_out/r02_dropin_train.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
