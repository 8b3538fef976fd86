/* generated model header of the reference build: nothing the node uses */
