import sys
sys.path.insert(0, 'scratch'); sys.path.insert(0, '.')
from proto2l import run
run('Octet', (16,16,16), 1, 2.0)
run('Octet', (16,16,16), 1, 4.0)
run('Octet', (24,24,24), 1, 3.0)
run('Octet', (24,24,24), 1, 4.0)
