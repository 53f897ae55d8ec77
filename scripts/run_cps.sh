# residency sweep of the step kernel on one box (MOOG_CTAS_PER_SM); CPS_LIST="3 4 5"
for r in ${CPS_LIST:-3 4 5 6}; do MOOG_CTAS_PER_SM=$r python bench.py --no-clocks --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('ctas/sm $r', 'value %.0f e2e %.0f step_ms %.3f'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms']))"; done
