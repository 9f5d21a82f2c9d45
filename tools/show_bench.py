import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(j["value"], j["e2e"]["value"], j["e2e_u8_ingest"]["value"])
