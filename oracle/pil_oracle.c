/* placeholder; renderer restatement follows */
