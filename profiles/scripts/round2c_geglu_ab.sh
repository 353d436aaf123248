for k in 6 4 8 12 16; do echo "== FF_GEGLU_CTAS_PER_SM=$k"; FF_GEGLU_CTAS_PER_SM=$k python profiles/gn_case.py 2>&1 | grep geglu; done
