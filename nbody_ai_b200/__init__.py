"""B200-native batched N-body TTV engine: drop-in for the generate()/analyze() path of
pearsonkyle/Nbody-AI.  See DESIGN.md and INTEGRATION.md."""
