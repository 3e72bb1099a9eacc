# final check of the committed library: the whole GPU suite, then smoke()
timeout 150 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/gputests_last.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_last.log 2>&1
tail -2 gpurun_out/gputests_last.log; tail -2 gpurun_out/smoke_last.log
