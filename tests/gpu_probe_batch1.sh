#!/bin/bash
# round-2 probe batch 1: localise the attention hang, get tracebacks of the failing tests, run the rest
P="timeout -s KILL 60 python tests/gpu_attn_probe.py"
echo "== attn fwd Lk=77 (SHORT)";            $P 2 8 256 77 64 64 fwd 2>&1 | tail -2; echo "rc=$?"
echo "== attn fwd Lk=128 (SHORT, no OOB)";   $P 2 8 256 128 64 64 fwd 2>&1 | tail -2; echo "rc=$?"
echo "== attn fwd Lk=77 NO_SHORT";           FDX_ATTN_NO_SHORT=1 $P 2 8 256 77 64 64 fwd 2>&1 | tail -2; echo "rc=$?"
echo "== attn bwd Lk=77";                    $P 2 8 256 77 64 64 bwd 2>&1 | tail -2; echo "rc=$?"
echo "== attn bwd Lk=128";                   $P 2 8 256 128 64 64 bwd 2>&1 | tail -2; echo "rc=$?"
echo "== attn fwd Lk=200 L=200 dh32";        $P 3 8 200 200 32 32 fwd 2>&1 | tail -2; echo "rc=$?"
echo "== attn bwd Lk=200 L=200 dh32";        $P 3 8 200 200 32 32 bwd 2>&1 | tail -2; echo "rc=$?"
echo "== attn fwd tiny";                     $P 2 8 4 4 64 64 fwd 2>&1 | tail -2; echo "rc=$?"
echo "== attn bwd tiny";                     $P 2 8 4 4 64 64 bwd 2>&1 | tail -2; echo "rc=$?"
export FDX_ATTN_UNFUSED=1
TAILN=25 PROBE_TIMEOUT=200 bash tests/gpu_probe_round2.sh \
  "tests/test_train_gpu.py::test_apply_gradients_then_apply_ema_equals_fused_step" \
  "tests/test_samplers_gpu.py::test_simple_ddpm_sampler_vs_oracle" \
  "tests/test_train_gpu.py::test_dynamic_scale_training_matches_unscaled" \
  "tests/test_train_gpu.py::test_ema_weights_stay_fresh_across_train_sample_train_sample" \
  "tests/test_train_gpu.py::test_sampler_graph_cache_is_bounded_and_tree_params_are_packed_once" \
  "tests/test_train_gpu.py::test_fit_runs_validation_sampling_from_ema_and_prefetches_uint8" \
  "tests/test_train_gpu.py::test_device_prefetcher_yields_device_batches_in_order" \
  "tests/test_train_gpu.py::test_train_step_parameters_and_ema_vs_oracle"
echo "=== test_unet_gpu.py (unfused attention)"
timeout -s KILL 500 python -m pytest tests/test_unet_gpu.py -q --tb=short 2>&1 | tail -40
