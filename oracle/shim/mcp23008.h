/* Oracle shim: the port-expander driver is not on the block path. */
