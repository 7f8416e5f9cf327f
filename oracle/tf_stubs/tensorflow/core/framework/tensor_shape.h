/* Empty stand-in so that the reference op sources compile without TensorFlow.
   Test infrastructure only (oracle/_ref build); see oracle/Makefile. */
