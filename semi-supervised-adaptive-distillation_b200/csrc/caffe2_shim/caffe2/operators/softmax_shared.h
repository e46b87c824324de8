// Present only so sources that include it (reference loss op does) still compile.
