timeout 600 python profiles/ab_bundle.py c5 2>&1 | cut -c1-400
