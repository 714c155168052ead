set -u
bash scripts/gpu_check.sh ref
bash scripts/profile_all.sh
