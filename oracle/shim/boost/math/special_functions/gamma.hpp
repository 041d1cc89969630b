// TEST INFRASTRUCTURE ONLY (oracle/): empty stand-in
