#!/bin/bash
tag=${1:-r02h}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q -k "boxqp or lims or back_pass or golden" > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $o/${tag}_pytest.log
tail -6 $o/${tag}_pytest.log | cut -c1-1500
timeout 600 python scripts/perf_lims.py 9472 3.0 > $o/${tag}_lims_loose.json 2> $o/${tag}_lims.err; cat $o/${tag}_lims_loose.json; tail -3 $o/${tag}_lims.err
timeout 600 python scripts/perf_lims.py 9472 0.3 > $o/${tag}_lims_tight.json 2>> $o/${tag}_lims.err; cat $o/${tag}_lims_tight.json
