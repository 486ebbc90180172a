for shape in "22 64 3" "22 128 1" "21 96 2" "19 200 3"; do
  set -- $shape
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --log-n $1 --cols $2 --rate-bits $3 2>&1 | grep -E '^\{|Error|error|failed' | head -2
done
