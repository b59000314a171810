# per-source-line executed instruction counts of a kernel from an ncu report + the current build
# usage: line_profile.sh <report.ncu-rep> <object.o> <mangled kernel> [units]
set -e
rep=$1; obj=$2; kern=$3; units=${4:-1e6}
tmp=$(mktemp -d)
ncu -i "$rep" --page source --csv --print-source sass > $tmp/sass.csv 2>/dev/null
obj=$(realpath $obj); (cd $tmp && cuobjdump -xelf all "$obj" > /dev/null && nvdisasm --print-line-info *.cubin > dis.txt)
python "$(dirname $0)/sass_by_line.py" $tmp/sass.csv $tmp/dis.txt "$kern" "$units"
