#!/bin/bash
# developer A/B: build variants of the library into e2e_parking_carla_b200/build/variants/<name>/ (run here), then
# tools/ab_run.sh on the GPU box times each of them.   usage: tools/ab_build.sh name "ENV=1 ENV2=2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p e2e_parking_carla_b200/build/variants
while [ $# -gt 0 ]; do
  name=$1; envs=$2; shift 2
  env $envs python -m e2e_parking_carla_b200.build --force >/dev/null 2>&1
  cp e2e_parking_carla_b200/libls_b200.so e2e_parking_carla_b200/build/variants/$name.so
  echo "built $name ($envs)"
done
python -m e2e_parking_carla_b200.build --force >/dev/null 2>&1
